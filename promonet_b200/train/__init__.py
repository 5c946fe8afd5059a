from . import ops
from .core import Trainer, train
from .evaluate import evaluate
