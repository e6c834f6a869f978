"""brapprox -- host-side mirror of LowRankApprox.jl's front-ends over libbrapprox.so.

The reference is Julia (absent from this image), so the host side above the
C ABI is Python here, mirroring the reference's entry points for the hot path
(same names, argument meaning and error behaviour):

    LRAOptions                     src/LowRankApprox.jl:77-148
    sketch(side, trans, A, order)  src/sketch.jl:35-50
    idfact / id                    src/id.jl:434-456
    pqrfact / pqr                  src/pqr.jl:285-320
    psvdfact / psvd / psvdvals     src/psvd.jl:238-308
    curfact / cur                  src/cur.jl:532-571 (index selection: two sketch-and-pivot passes)
    pheigfact / pheig / pheigvals  src/pheig.jl:276-319
    snorm / snormdiff              src/snorm.jl:14-53
    prange                         src/prange.jl:14-62

The product path is the CUDA library only: importing this package without a
loadable libbrapprox.so raises, and every call fails loudly (BraError) when no
B200 is usable.  Nothing here imports ``oracle/``.
"""
from ._binding import (  # noqa: F401
    BraError,
    CURPackedU,
    Context,
    DeviceMatrix,
    IDPackedV,
    LRAOptions,
    PartialHermEigen,
    PartialQR,
    PQRFactors,
    PartialSVD,
    SKETCH_CODES,
    lib,
    lib_path,
)
from ._dist import block_shard, broadcast_bytes, init_comm, row_shard  # noqa: F401
from ._frontend import (  # noqa: F401
    CUR,
    cur,
    curfact,
    default_context,
    geqp3_adap,
    id,
    idfact,
    idfact_batched,
    idfact_batched_device,
    idfact_device,
    pheig,
    pheigfact,
    pheigvals,
    pqr,
    pqrfact,
    pqrfact_device,
    psvd,
    psvdfact,
    psvdfact_device,
    psvdvals,
    probe_exchange_latency,
    probe_fp64_peak,
    sketch,
    sketchfact,
    prange,
    snorm,
    snormdiff,
    trsolve_T,
)
