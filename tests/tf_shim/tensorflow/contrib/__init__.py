from . import framework
