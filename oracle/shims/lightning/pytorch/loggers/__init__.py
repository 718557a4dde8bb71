"""ORACLE / TEST INFRASTRUCTURE ONLY."""


class WandbLogger:
    def __init__(self, *a, **k):
        pass
