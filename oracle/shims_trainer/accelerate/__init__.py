"""ORACLE / TEST INFRASTRUCTURE ONLY — single-process stand-in for the slice of HF accelerate used by
training_duett/trainer.py:12-13,217-218,418-419."""
import contextlib
import torch


class DistributedDataParallelKwargs:
    def __init__(self, **kw):
        self.kw = kw


class Accelerator:
    def __init__(self, mixed_precision="no", kwargs_handlers=None, **kw):
        self.mixed_precision = mixed_precision
        self.device = torch.device("cpu")
        self.is_main_process = True
        self.num_processes = 1
        self.process_index = 0

    def print(self, *a, **k):
        print(*a, **k)

    def prepare(self, *objs):
        return objs if len(objs) != 1 else objs[0]

    def backward(self, loss):
        loss.backward()

    def unwrap_model(self, m):
        return m

    def autocast(self):
        if self.mixed_precision == "bf16":
            return torch.autocast(self.device.type, dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def wait_for_everyone(self):
        pass

    def reduce(self, t, reduction="sum"):
        return t
