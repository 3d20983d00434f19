"""eventpretrain_b200 — B200-native (sm_100a) input hot path of BIT-Vision/EventPretrain.

Raw events -> dense tensors -> diff-map target -> mask / patchify / visible-token gather, as
hand-written CUDA kernels behind a C ABI (include/eventpretrain_b200.h), with drop-in Python
functions that keep the reference's call signatures.  No CPU fallback.
"""
from . import config  # noqa: F401
from ._lib import LIB_PATH, NativeLibraryMissing, load as load_library  # noqa: F401
from .augmentation import (add_noise_events, erase_and_add_events, events_augment, events_reshape,  # noqa: F401
                           get_random_index, reshape_scale)
from .dataset_utils import (events_to_EvRep, events_to_image_ecdp, events_to_image_mem,  # noqa: F401
                            events_to_voxel_grid, remove_hot_pixel_mem)
from .events import (BadEventsError, EventCollator, RaggedEvents, bin_events, bin_events_aos, evrep, from_soa,  # noqa: F401
                     collate_events, collate_transport, mem_hotpixel, normalise, pack_events, time_surface)
from .masking import (block_mask_expand, convvit_fuse_stages, convvit_keep_masks, gather_tokens, gather_tokens_nchw, len_keep_of,  # noqa: F401
                      mask_from_noise, patch_density, random_masking, swin_apply_mask, swin_scatter_dense, unshuffle_tokens)
from .pipeline import MaskedInputPipeline  # noqa: F401
from .swin_grouping import GroupingModule, group_windows, knapsack, patch_merging_order  # noqa: F401
from .readers import pack_soa, read_ddd17_memmap, read_nimagenet_npz  # noqa: F401
from .reshape import (diffmap_frames, frame2emb, patchify_gather, reconstruct_loss, target_normpix,  # noqa: F401
                      target_patch_loss)

from .view_augment import (ViewChoice, apply_views, prepare_views, draw_crop, draw_evg_choice, draw_frame_choice,  # noqa: F401
                           evg_augment, frame_augment)

__version__ = "0.1.0"
