"""gesture2vec_b200 -- the vector-quantizer hot path of Gesture2Vec on NVIDIA B200 (sm_100a).

Only what the path needs: the C-ABI CUDA library (csrc/, include/g2v_vq.h), its ctypes binding,
and the host-side mirror of the reference's quantizer modules.
"""
from . import _lib  # noqa: F401
from .functional import (prepare_codebook, vq_search, vq_apply, quantize, tokenize, tokenize_host, pinned_empty,  # noqa: F401
                         one_hot, stats_finalize, ema_update, packed_numel, step_finalize, vq_search_exact,
                         vq_search_wide, pad_rows, gemm, linear)
from .quantizers import (DAE_VQ_Payam, DAE_VQ_Payam_EMA, VQVAE_VQ_Payam, VQVAE_VQ_Payam_EMA,  # noqa: F401
                         VectorQuantizerEMA, VQVAE_VQ_Payam_GSSoft, FLAVOURS)
from .reference_patch import patch_reference, unpatch_reference, swap_vq_layer  # noqa: F401
from .distributed import (shard_rows, StatsAllReduce, enable_data_parallel_ema, packed_layout,  # noqa: F401
                          broadcast_quantizer_state)

from .soft import soft_quantize  # noqa: F401
from . import functional  # noqa: F401
from .kmeans import KMeans, kmeans_update  # noqa: F401
from .tokenizer import GestureTokenizer, chunk_rows_from_hidden  # noqa: F401

__version__ = "0.2.0"
