"""headnerf-b200: sm_100a implementation of NeRF-3DTalker's HeadNeRF rendering hot path behind the
reference's own module interface (HeadNeRFNet.forward, BaseOptions, the .pth state-dict layout).
The directory name contains hyphens; import it with importlib.import_module("nerf-3dtalker-code_b200")."""
from .options import BaseOptions
from .neural_renderer import NeuralRenderer, PixelShuffleUpsample, Blur
from .headnerf_net import HeadNeRFNet, MLPforNeRF
from . import ops
from . import dist
from . import train
from . import sampling
from .sampling import FineSample
from .train import HeadNeRFLossUtils, FusedAdam, Audio2style, save_checkpoint, load_checkpoint
from . import _lib
from .build import build as build_library

__all__ = ["BaseOptions", "NeuralRenderer", "PixelShuffleUpsample", "Blur", "HeadNeRFNet", "MLPforNeRF",
           "ops", "build_library", "train", "HeadNeRFLossUtils", "FusedAdam", "Audio2style", "save_checkpoint", "load_checkpoint", "FineSample", "sampling"]
