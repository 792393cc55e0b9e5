from .networks import ConditionalVNet, SharedVNet  # noqa: F401
from .pretrain import (CoupledResidualFM, FlowMatching, FusedAdam, _rademacher, score_pair, score_pair_host,  # noqa: F401
                       set_default_precision, train_step)
