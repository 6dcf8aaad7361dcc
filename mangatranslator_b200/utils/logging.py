"""log_message with the reference's signature (utils/logging.py:6-19): a locked print gated by verbose/always_print."""
import threading

_lock = threading.Lock()


def log_message(message, verbose=False, always_print=False):
    if verbose or always_print:
        with _lock:
            print(str(message))
