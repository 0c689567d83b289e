"""Model registry - drop-in for reference model_dict.py:4-11: `get_model(args)` returns the *module* whose `.Model` the caller
instantiates; `args.model` is one of the four keys the reference registers (anything else is a KeyError, as there)."""
import importlib

_KEYS = ("Transolver_Irregular_Mesh", "Transolver_Structured_Mesh_2D", "Transolver_Structured_Mesh_3D",
         "Transolver_Structured_Mesh2D_Encoder")


def get_model(args):
    if args.model not in _KEYS:
        raise KeyError(args.model)
    return importlib.import_module(f"{__package__}.model.{args.model}")
