"""Import the real reference hot path from /root/reference — TEST INFRASTRUCTURE, build container only.

/root/reference does not exist on the GPU box; everything that runs there uses
oracle/headnerf_oracle.py (pinned against this import by tests/test_oracle_vs_reference.py) and the
committed fixtures under tests/golden/. Recipe from SURVEY.md Appendix B: put the reference root and
a kornia.filters.filter2d shim on sys.path, reset sys.argv (HeadNeRFOptions.py:55 runs argparse at
import time), import NetWorks.HeadNeRFNet."""
import os
import sys

REFERENCE_ROOT = os.environ.get("HN_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "NetWorks", "HeadNeRFNet.py"))


def load():
    """-> (BaseOptions, HeadNeRFNet) classes of the unmodified reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    argv, sys.argv = sys.argv, [sys.argv[0] if sys.argv else "x"]
    try:
        from HeadNeRFOptions import BaseOptions          # noqa
        from NetWorks.HeadNeRFNet import HeadNeRFNet     # noqa
    finally:
        sys.argv = argv
    return BaseOptions, HeadNeRFNet


def build(featmap_size: int, pred_img_size: int, featmap_nc: int = 256, hidden: int = None, n_samples: int = None):
    """Instantiate the reference net the way the trainer does (talker_trainer.py:693-699)."""
    BaseOptions, HeadNeRFNet = load()
    opt = BaseOptions({"featmap_size": featmap_size, "featmap_nc": featmap_nc, "pred_img_size": pred_img_size})
    if hidden is not None:
        opt.mlp_hidden_nchannels = hidden
    if n_samples is not None:
        opt.num_sample_coarse = n_samples
    net = HeadNeRFNet(opt, include_vd=False, hier_sampling=False)
    return opt, net
