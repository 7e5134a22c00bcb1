def imread(*a, **k):
    raise RuntimeError("imageio is not available offline")
