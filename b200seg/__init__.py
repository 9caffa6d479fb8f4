"""Import alias: the product package lives in `general-medical-image-segmentation-cnn-framework_b200/` (a directory
name Python cannot import directly).  `import b200seg` resolves its sub-modules there."""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))
PACKAGE_DIR = _os.path.join(_os.path.dirname(_HERE), "general-medical-image-segmentation-cnn-framework_b200")
__path__ = [PACKAGE_DIR]
with open(_os.path.join(PACKAGE_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(PACKAGE_DIR, "__init__.py"), "exec"))
