"""tests/test_gpu_multi.py once more, with the exchange of the multi-device layer going through pinned host memory (the library's default
beyond two devices since the 8-GPU run of round 2; csrc/upg_multi.cu peer_exchange()).  Needs >= 2 GPUs: everything skips on a 1-GPU box.
(First run on hardware is the driver's: tests/test_gpu_unverified.py.)"""
import os

os.environ["UPG_MULTI_EXCHANGE"] = "host"

from test_gpu_multi import *  # noqa: E402,F401,F403
