from .loss import NCELearnableTempLoss, build_loss_func  # noqa: F401
from .adamw import AdamW, build_e2e_optimizer_w_lr_mul, clip_grad_norm_, get_lr_sched, setup_e2e_optimizer  # noqa: F401
