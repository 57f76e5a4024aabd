from .stereo_focal_loss import StereoFocalLoss  # noqa: F401
