from .fargan import FarganGenerator
from .generator import Generator
from . import init
