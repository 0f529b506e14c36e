from .clip_vip import CLIPModel, ClipVipConfig, TowerConfig  # noqa: F401
from .vidclip import VidCLIP  # noqa: F401
