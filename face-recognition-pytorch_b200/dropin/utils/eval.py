# replaces utils/eval.py of the reference
from face_recognition_pytorch_b200.eval import pair_score, cross_score, performance_roc, performance_acc, kfold_accuracy  # noqa: F401
