"""str2bool / setup_seed helpers of the reference's utils.py:9-35 (rewritten: distutils is gone in 3.12)."""
import os
import random

import numpy as np
import torch


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = str(v).strip().lower()
    if s in ("y", "yes", "t", "true", "on", "1"):
        return True
    if s in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError("invalid truth value %r" % (v,))


def setup_seed(random_seed, cudnn_deterministic=True):
    """utils.py:13-35 seeds random / numpy / PYTHONHASHSEED (torch.manual_seed is commented out there,
    utils.py:26-27); the copy in resnet.py:110-120 also seeds torch.  We seed torch too so that parameter
    initialisation is reproducible."""
    random.seed(random_seed)
    np.random.seed(random_seed)
    os.environ["PYTHONHASHSEED"] = str(random_seed)
    torch.manual_seed(random_seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(random_seed)
