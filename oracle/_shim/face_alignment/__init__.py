"""Import stub (test infrastructure): Utils/HeadNeRFLossUtils.py imports face_alignment at module level but only its dead
calc_disp_loss touches it; the package is not installed here."""


class LandmarksType:
    _2D = 0


class FaceAlignment:
    def __init__(self, *a, **k):
        raise RuntimeError("face_alignment is a stub in this container")
