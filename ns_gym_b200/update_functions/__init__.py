from .distribution import *  # noqa: F401,F403
from .single_param import *  # noqa: F401,F403
from . import distribution, single_param

__all__ = [*distribution.__all__, *single_param.__all__]
