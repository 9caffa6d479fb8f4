"""Hydra-style configuration without Hydra (not installed in the target image): `python train.py config=unet
config.batch_size=2 config.patch_size=128,128,128` composes conf/config.yaml with conf/config/<name>.yaml and applies the
dotted overrides, giving the same `config` namespace the reference's `@hydra.main` hands to `main` (train.py:310-312)."""
import os

import yaml

CONF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")


class Config(dict):
    """dict with attribute access (what the reference code does with its DictConfig)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _parse_value(text):
    try:
        return yaml.safe_load(text)
    except yaml.YAMLError:
        return text


def _tuple_of_ints(v):
    if isinstance(v, str) and "," in v:
        return tuple(int(t) for t in v.split(","))
    if isinstance(v, (list, tuple)):
        return tuple(int(t) for t in v)
    return v


def compose(argv=(), conf_dir=CONF_DIR):
    base = yaml.safe_load(open(os.path.join(conf_dir, "config.yaml")))
    group = "unet"
    for entry in base.get("defaults", []):
        if isinstance(entry, dict) and "config" in entry:
            group = entry["config"]
    overrides = []
    for arg in argv:
        if "=" not in arg:
            raise ValueError("expected key=value, got %r" % arg)
        k, v = arg.split("=", 1)
        if k == "config":
            group = v
        else:
            overrides.append((k, v))
    path = os.path.join(conf_dir, "config", group + ".yaml")
    if not os.path.exists(path):
        raise FileNotFoundError("no config group %r (looked for %s)" % (group, path))
    cfg = Config(base.get("config", {}))
    cfg.update(yaml.safe_load(open(path)))
    for k, v in overrides:
        if not k.startswith("config."):
            raise ValueError("only config.<key> overrides are supported, got %r" % k)
        cfg[k[len("config."):]] = _parse_value(v)
    for k in ("patch_size", "crop_or_pad_size", "patch_overlap", "volume_size"):
        if k in cfg:
            cfg[k] = _tuple_of_ints(cfg[k])
    if cfg.get("ckpt") in ("None", "none", ""):
        cfg["ckpt"] = None
    cfg.setdefault("hydra_path", cfg.get("output_dir", "./logs"))
    return cfg


def build_model(config):
    """The `config.network` switch of train.py:324-373 / predict.py:233-276 for the models on the b200seg path."""
    net = config.network
    if net == "unet":
        from .models.three_d.unet3d import UNet3D
        return UNet3D(in_channels=config.in_classes, out_channels=config.out_classes, init_features=32)
    if net == "res_unet":
        from .models.three_d.residual_unet3d import UNet
        return UNet(in_channels=config.in_classes, n_classes=config.out_classes, base_n_filter=32)
    if net == "vnet":
        from .models.three_d.vnet3d import VNet
        return VNet(in_channels=config.in_classes, classes=config.out_classes)
    if net == "densevoxelnet":
        from .models.three_d.densevoxelnet3d import DenseVoxelNet
        return DenseVoxelNet(in_channels=config.in_classes, classes=config.out_classes)
    if net == "highresnet":   # not wired into the reference's train.py; BASELINE.json config 4 names it
        from .models.three_d.highresnet import HighRes3DNet
        return HighRes3DNet(config.in_classes, config.out_classes)
    if net == "csrnet":       # train.py:366-369 (init_features defaults to 64 there)
        from .models.three_d.csrnet import CSRNet
        return CSRNet(in_channels=config.in_classes, out_channels=config.out_classes)
    if net == "er_net":       # train.py:332-335
        from .models.three_d.ER_net import ER_Net
        return ER_Net(classes=config.out_classes, channels=config.in_classes)
    if net == "re_net":       # train.py:336-339
        from .models.three_d.RE_net import RE_Net
        return RE_Net()
    if net == "dunet":  # train.py:370-373
        from .models.three_d.Double_Unet import Double_Unet
        return Double_Unet(in_channels=config.in_classes, out_channels=config.out_classes)
    raise ValueError("network %r is not on the b200seg path (supported: unet, res_unet, vnet, densevoxelnet, "
                     "highresnet, csrnet, er_net, re_net, dunet)" % net)


def weights_init_normal(init_type):
    """train.py:33-61, name-based like the reference: every module whose class name contains Conv / Linear and has a
    weight (ConvTranspose3d included) is re-initialised, its bias zeroed; BatchNorm3d / InstanceNorm / PReLU untouched."""
    import torch

    def init_func(m):
        classname = m.__class__.__name__
        gain = 0.02
        if classname.find("BatchNorm2d") != -1:
            if hasattr(m, "weight") and m.weight is not None:
                torch.nn.init.normal_(m.weight.data, 1.0, gain)
            if hasattr(m, "bias") and m.bias is not None:
                torch.nn.init.constant_(m.bias.data, 0.0)
        elif hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                torch.nn.init.normal_(m.weight.data, 0.0, gain)
            elif init_type == "xavier":
                torch.nn.init.xavier_normal_(m.weight.data, gain=gain)
            elif init_type == "xavier_uniform":
                torch.nn.init.xavier_uniform_(m.weight.data, gain=1.0)
            elif init_type == "kaiming":
                torch.nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                torch.nn.init.orthogonal_(m.weight.data, gain=gain)
            elif init_type == "none":
                m.reset_parameters()
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                torch.nn.init.constant_(m.bias.data, 0.0)
    return init_func
