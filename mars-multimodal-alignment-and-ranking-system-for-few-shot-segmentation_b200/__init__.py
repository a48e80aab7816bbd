"""marsb200: B200-native (sm_100a) proposal scoring / ranking / merging stage of MARS.

Host side in Python mirroring the reference's classes; the arithmetic lives in
libmarsb200.so (hand-written CUDA behind the C ABI of include/marsb200.h).
Importing this package loads the shared library and fails if it is missing -
there is no CPU fallback.
"""
from . import _lib, ops, synthetic  # noqa: F401  (loads libmarsb200.so)
from ._lib import MarsB200Error  # noqa: F401
from .episodes import (InterleavedRanking, PipelinedRanking, RankingConfig, RankingEngine, decode_records, gather_records,  # noqa: F401
                       kernel_launches_per_run, shard_range)
from .synthetic import CONFIGS, EpisodeShape, make_episode, masks_to_rle, stack_episodes, to_device  # noqa: F401
from .components import (FilteringMergingModule, PriorInformationRefinementModule,  # noqa: F401
                         VisualVisualAlignmentModule)
from .MARS import MARS, build_MARS_fss  # noqa: F401
from .matcher_scoring import MatcherScorer, PatchMatcher  # noqa: F401
from .Matcher import Matcher, RobustPromptSampler  # noqa: F401
from .evaluation import AverageMeter, Evaluator  # noqa: F401
from .host_ingest import HostMaskIngest  # noqa: F401

__version__ = "0.1.0"
