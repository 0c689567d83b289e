"""Model registry - drop-in for reference model_dict.py:4-11: `get_model(args)` returns the *module* whose `.Model`
the caller instantiates.  The 3D entry is outside this round's scope."""
from .model import Transolver_Irregular_Mesh, Transolver_Structured_Mesh_2D, Transolver_Structured_Mesh2D_Encoder

_NOT_BUILT = ('Transolver_Structured_Mesh_3D',)


def get_model(args):
    model_dict = {
        'Transolver_Irregular_Mesh': Transolver_Irregular_Mesh,
        'Transolver_Structured_Mesh_2D': Transolver_Structured_Mesh_2D,
        'Transolver_Structured_Mesh2D_Encoder': Transolver_Structured_Mesh2D_Encoder,
    }
    if args.model in _NOT_BUILT:
        raise NotImplementedError(f"{args.model} is registered by the reference but not on the B200 hot path yet (SURVEY.md §8f)")
    return model_dict[args.model]
