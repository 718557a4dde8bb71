"""ORACLE / TEST INFRASTRUCTURE ONLY."""
from .. import Callback


class ModelCheckpoint(Callback):
    def __init__(self, *a, **k):
        self.best_model_path = ""
