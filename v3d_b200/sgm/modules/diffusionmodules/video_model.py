from v3d_b200.unet import VideoUNet  # noqa: F401  (reference: sgm/modules/diffusionmodules/video_model.py:84)
