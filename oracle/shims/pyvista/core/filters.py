def _get_output(*a, **k):
    raise NotImplementedError("pyvista stand-in: plotting is out of scope")


def _update_alg(*a, **k):
    raise NotImplementedError("pyvista stand-in: plotting is out of scope")
