"""Error types mirroring ``PoastaError`` (/root/reference/src/errors.rs:10-40) at the C-ABI status codes."""


class PoastaError(Exception):
    pass


class WeightsUnequalSize(PoastaError):
    pass


class InvalidAlignment(PoastaError):
    pass


class GraphError(PoastaError):
    pass


class CudaError(PoastaError):
    """No device / kernel failure. There is no CPU fallback behind the aligner."""


class NoAlignment(PoastaError):
    """The reference panics here ("Could not align sequence!", aligner/astar.rs:143)."""


class Unsupported(PoastaError):
    pass


_BY_CODE = {-1: CudaError, -2: ValueError, -3: WeightsUnequalSize, -4: InvalidAlignment, -5: GraphError,
            -6: NoAlignment, -7: Unsupported, -8: IOError, -9: PoastaError}


def check(code, what=""):
    if code == 0:
        return
    raise _BY_CODE.get(code, PoastaError)(f"{what}: poasta_status {code}")
