"""stub: the reference's sglang/utils.py imports IPython.display at module import (notebook helpers only)."""
