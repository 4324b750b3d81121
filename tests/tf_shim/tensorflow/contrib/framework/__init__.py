from . import python
