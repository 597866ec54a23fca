import time


def utc_epoch_now() -> int:
    """Milliseconds since the epoch (reference src/dnet/utils/time.py)."""
    return int(time.time() * 1000)
