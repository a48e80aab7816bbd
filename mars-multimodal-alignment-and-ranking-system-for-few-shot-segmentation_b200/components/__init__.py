from .PriorInformationRefinementModule import PriorInformationRefinementModule
from .VisualVisualAlignmentModule import VisualVisualAlignmentModule
from .FilteringMergingModule import FilteringMergingModule

__all__ = ["PriorInformationRefinementModule", "VisualVisualAlignmentModule", "FilteringMergingModule"]
