# replaces nets/ArcFace.py of the reference
from face_recognition_pytorch_b200.arcface import ArcFace, CosFace, CombinedMarginLoss  # noqa: F401
