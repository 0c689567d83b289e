"""Process-wide default operand precision of the large contractions ("fp32" | "bf16" | "fp32_exact").

"fp32": fp32 operands as 3xTF32 split products on the tensor cores (error <= 2^-21 per product, fp32 accumulate);
"fp32_exact": fp32 FMA on the SIMT engine; "bf16": bf16 operands, fp32 accumulate."""
import os

_DEFAULT = os.environ.get("TBNS_PRECISION", "bf16")


def set_default_precision(p: str) -> None:
    global _DEFAULT
    if p not in ("fp32", "bf16", "fp32_exact"):
        raise ValueError("precision must be 'fp32', 'fp32_exact' or 'bf16'")
    _DEFAULT = p


def get_default_precision() -> str:
    return _DEFAULT
