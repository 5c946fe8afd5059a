"""Seeded synthetic inputs shared by the oracle, the tests and bench.py: the
generators live with the product (promonet_b200/synthetic.py) so that bench.py's
B200 arm does not import this package; re-exported here for the tests."""
from promonet_b200.synthetic import audio, synthesis  # noqa: F401
