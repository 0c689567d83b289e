"""Model registry - drop-in for reference model_dict.py:4-11: `get_model(args)` returns the *module* whose `.Model`
the caller instantiates (all four reference keys)."""
from .model import (Transolver_Irregular_Mesh, Transolver_Structured_Mesh_2D, Transolver_Structured_Mesh_3D,
                    Transolver_Structured_Mesh2D_Encoder)


def get_model(args):
    model_dict = {
        'Transolver_Irregular_Mesh': Transolver_Irregular_Mesh,
        'Transolver_Structured_Mesh_2D': Transolver_Structured_Mesh_2D,
        'Transolver_Structured_Mesh_3D': Transolver_Structured_Mesh_3D,
        'Transolver_Structured_Mesh2D_Encoder': Transolver_Structured_Mesh2D_Encoder,
    }
    return model_dict[args.model]
