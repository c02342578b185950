"""Import alias: the package directory is named `face-recognition-pytorch_b200/` (not a valid Python identifier),
so this shim maps `import face_recognition_pytorch_b200` onto it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "face-recognition-pytorch_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
