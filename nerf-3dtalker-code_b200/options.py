"""BaseOptions — same attribute names and defaults as the reference's HeadNeRFOptions.BaseOptions
(HeadNeRFOptions.py:5-34), so checkpoints' "para" dicts and callers' option objects work unchanged."""


class BaseOptions(object):
    def __init__(self, para_dict=None) -> None:
        super().__init__()
        self.bg_type = "white"          # "white" or "black" background feature map
        self.iden_code_dims = 100
        self.expr_code_dims = 79
        self.text_code_dims = 100
        self.illu_code_dims = 27
        self.auxi_shape_code_dims = 179
        self.auxi_appea_code_dims = 127
        self.num_sample_coarse = 64
        self.num_sample_fine = 128
        self.world_z1 = 2.5
        self.world_z2 = -3.5
        self.mlp_hidden_nchannels = 384
        para = para_dict or {}
        self.featmap_size = para.get("featmap_size", 32)
        self.featmap_nc = para.get("featmap_nc", 256)
        self.pred_img_size = para.get("pred_img_size", 256)
