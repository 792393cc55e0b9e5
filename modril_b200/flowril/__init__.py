from .networks import ConditionalVNet, GradNorm2D, SharedVNet  # noqa: F401
from .pretrain import (CoupledResidualFM, FlowMatching, FusedAdam, _rademacher, score_pair, score_pair_host,  # noqa: F401
                       reg_losses_native, set_default_precision, set_train_precision, tensor_grad_norm_sum, train_step, TrainStepGraph)
