from .classify import HarNetRoIHead, RoIAlign, RoIPool  # noqa: F401
from .frcnn import FasterRCNN  # noqa: F401
from .frcnn_training import AnchorTargetCreator, FasterRCNNTrainer, ProposalTargetCreator  # noqa: F401
from .rpn import ProposalCreator, RegionProposalNetwork  # noqa: F401
