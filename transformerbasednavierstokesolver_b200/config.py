"""Process-wide default operand precision of the large contractions ("fp32" | "bf16")."""
import os

_DEFAULT = os.environ.get("TBNS_PRECISION", "bf16")


def set_default_precision(p: str) -> None:
    global _DEFAULT
    if p not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _DEFAULT = p


def get_default_precision() -> str:
    return _DEFAULT
