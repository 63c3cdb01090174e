"""CPU oracle for the multi-codebook hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (quantization_b200/) never does.
"""
from .mcq_oracle import (  # noqa: F401
    OracleError,
    build,
    compute_indexes,
    decode,
    encode,
    max_threads,
    pack,
    unpack,
)
