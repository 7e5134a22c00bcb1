class LPIPS:
    def __init__(self, *a, **k):
        raise RuntimeError("lpips is not available offline")
