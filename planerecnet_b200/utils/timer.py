"""Section timer with the reference's interface (utils/timer.py: env / reset / total_time /
print_stats / enable_all / disable_all), measured with CUDA events on the current stream.  Disabled by
default so the forward issues no synchronisation; eval-style callers enable it explicitly."""
import torch

_enabled = False
_times = {}
_order = []


class env:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _enabled:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if _enabled:
            self.end.record()
            self.end.synchronize()   # like utils/timer.py:101-103
            if self.name not in _times:
                _order.append(self.name)
            _times[self.name] = _times.get(self.name, 0.0) + self.start.elapsed_time(self.end)
        return False


def enable_all():
    global _enabled
    _enabled = True


def disable_all():
    global _enabled
    _enabled = False


def reset():
    _times.clear()
    _order.clear()


def total_time():
    return sum(_times.values())


def print_stats():
    print()
    width = max([len(k) for k in _order] + [4])
    print(f"{'Name':>{width}} | Time (ms)")
    print("-" * (width + 13))
    for k in _order:
        print(f"{k:>{width}} | {_times[k]:9.4f}")
    print("-" * (width + 13))
    print(f"{'Total':>{width}} | {total_time():9.4f}")
    print()
