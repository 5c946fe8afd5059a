from .generator import Generator
from . import init
