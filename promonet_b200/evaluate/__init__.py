from .metrics import Metrics
