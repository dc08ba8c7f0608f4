"""B200-native SEGNN message-passing hot path (drop-in for model_type=segnn / dataloader_type=segnn_nbody).

Importing this package loads ``libsegnn_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
include/segnn_b200.h) and fails loudly if it is missing: there is no CPU, PyTorch or Triton fallback."""
from . import _lib, ops, packing  # noqa: F401
from . import macros, simulator  # noqa: F401,E402
from .graph import GraphBatch, build_graph_with_knn  # noqa: F401
from .irreps import Irreps, weight_balanced_irreps  # noqa: F401
from .o3_building_blocks import (BatchNorm, InstanceNorm, O3TensorProduct, O3TensorProductSwishGate,  # noqa: F401
                                 O3Transform)  # noqa: F401
from .segnn import SEGNN, SEGNNLayer  # noqa: F401

WeightBalancedIrreps = weight_balanced_irreps
from .rollout import SelfFeedRollout, run_inference, shard_simulations  # noqa: F401,E402
from .dataloader import (GravityDatasetOtf, NBodySystemDataset, SegnnNBodyDataLoader,  # noqa: F401,E402
                         SegnnNbodyOfflineDataloader)
from .trainer import TrainStep, allreduce_gradients, noam_rate, target_common_loss  # noqa: F401,E402
from . import checkpoint, torch_ops  # noqa: F401,E402  (torch_ops registers torch.ops.segnn_b200.*)
from .checkpoint import load_checkpoint, load_model_for_inference, save_model  # noqa: F401,E402
