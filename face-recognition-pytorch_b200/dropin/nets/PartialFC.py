# replaces nets/PartialFC.py of the reference
from face_recognition_pytorch_b200.partial_fc import PartialFC, PartialFCAdamW  # noqa: F401
from face_recognition_pytorch_b200.arcface import ArcFace  # noqa: F401
