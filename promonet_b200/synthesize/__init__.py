from .core import *
