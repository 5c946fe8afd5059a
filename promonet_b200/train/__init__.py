from . import ops
from .core import Trainer, train
