def _ops():
    """Late import so that `import tgis_b200.utils.*` works on a CPU box; calling an op without CUDA raises."""
    from .. import ops
    return ops
