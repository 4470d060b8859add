"""Alias package: `import dmesh_renderer` resolves to the B200-native implementation, so existing DMesh code
(`from dmesh_renderer import TriRenderer, TriRenderSettings, ...`) runs unchanged with this repository on the
path instead of the reference's package (reference surface: dmesh_renderer/__init__.py:13-488)."""
from dmesh_renderer_b200 import *  # noqa: F401,F403
from dmesh_renderer_b200 import _C, __all__, set_deterministic  # noqa: F401
