"""Import shim: the package directory is ``ai-camera_b200/`` (a hyphen cannot be imported),
so ``import ai_camera_b200`` resolves to this module, which turns itself into that package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "ai-camera_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
