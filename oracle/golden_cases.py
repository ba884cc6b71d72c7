"""Case list and helpers shared by oracle/make_goldens.py and tests (TEST INFRASTRUCTURE ONLY)."""
import numpy as np
import torch

DEFORM_CASES = [
    dict(name="deform1d_n193_b2", b=2, n=193, seed=11),
    dict(name="deform1d_n128_b1_even", b=1, n=128, seed=12),
    dict(name="deform1d_n517_b1", b=1, n=517, seed=13),
]
NYSTROM_CASES = [
    dict(name="nystrom_d64_m16_n100_b2", b=2, n=100, dim=64, dim_head=8, m=16, seed=21),
    dict(name="nystrom_d512_m256_n300_b1", b=1, n=300, dim=512, dim_head=64, m=256, seed=22),
    dict(name="nystrom_d128_m64_n256_b1_nopad", b=1, n=256, dim=128, dim_head=16, m=64, seed=23),
    # B = 2 at head sizes the GPU library supports (H * d >= 128): pins the cross-bag coupling of the pinv init (T3)
    dict(name="nystrom_d128_m64_n200_b2", b=2, n=200, dim=128, dim_head=16, m=64, seed=24),
    # CMTA's Nystrom configuration (cmta_utils.py:858-874: dim 256, dim_head 32, 128 landmarks), B = 2, front pad 84
    dict(name="nystrom_d256_m128_n300_b2_cmta", b=2, n=300, dim=256, dim_head=32, m=128, seed=25),
]
TOWER_CASES = [dict(name="dctmil_n301_b1", B=1, N=300, seed=31)]
TRANSMIL_CASES = [
    dict(name="transmil_n150_b1", B=1, N=150, seed=41),
    # BASELINE.json config 1 (N = 6 000: grid 78^2, n = 6 085, front pad 59, l = 24) and the 16k bag (pad 255, l = 65):
    # the reference itself runs these in seconds on the CPU, so the goldens come from it, not from the oracle
    dict(name="transmil_n6000_b1", B=1, N=6000, seed=42),
    dict(name="transmil_n16384_b1", B=1, N=16384, seed=43),
]
PATHOMIC_CASES = [
    dict(name="pathomic_diag_n260_b2", B=2, N=260, seed=51, task="diag2021"),
    dict(name="pathomic_surv_n132_b1", B=1, N=132, seed=52, task="survival"),
]
# raw-score co-attention (models/MultiheadAttention.py) as MCAT / CMTA call it: 1 head, E = 256.  L queries over S keys.
COATTN_CASES = [
    dict(name="coattn_mcat_l4_s300_b2", B=2, L=4, S=300, seed=61),          # model.py:1047 (4 genomic queries over the patches)
    dict(name="coattn_cmta_l333_s6_b1", B=1, L=333, S=6, seed=62),          # model.py:1229-1233 (patches ask, genomic keys)
    dict(name="coattn_cmta_l6_s333_b1", B=1, L=6, S=333, seed=63),          # model.py:1234-1238 (and back)
]
# utils/loss.py batch losses at world_size 1 (N = batch): attention maps [N, 8, L1, L2], omic [N, 128], vgrid [8 N, 2, 12, 12]
LOSS_CASES = [
    dict(name="losses_n4_l50x12", N=4, L1=50, L2=12, seed=81),
    dict(name="losses_n8_l37x20", N=8, L1=37, L2=20, seed=82),
]
# DeformCrossAttention2D as the teacher / student encoders build it (models/Modules.py:107-126: dim 128, 8 heads = 8 offset
# groups, dim_head 64, stride 4, offset_scale 4) on side x side grids; s50 is the reference's own bag (2 500 patches, 144 keys)
DEFORM2D_CASES = [
    dict(name="deform2d_s20_b2", b=2, side=20, seed=91),
    dict(name="deform2d_s50_b1", b=1, side=50, seed=92),
    dict(name="deform2d_s23_b1", b=1, side=23, seed=93),      # 25 keys: not a multiple of the 16-key tile
]
# ClusterMergeNet (models/ClusterMergeNet.py:183-207) at sample ratios that give 8 and 32 clusters (Modules.py:258-261)
CLUSTER_CASES = [
    dict(name="clustermerge_n400_b2", B=2, N=400, ratio=0.02, seed=95),
    dict(name="clustermerge_n2500_b1", B=1, N=2500, ratio=0.0128, seed=96),
]
# the teacher / student callers of the 2-D operator (models/Modules.py: TeacherNet :357-397, StudentNet :429-458), eval mode
TEACHER_CASES = [
    dict(name="teacher_s20_b2", kind="teacher", B=2, side=20, seed=101),
    dict(name="teacher_s50_b1", kind="teacher", B=1, side=50, seed=102),
    dict(name="student_s50_b2", kind="student", B=2, side=50, seed=103),       # 2 500 tokens -> ceil(0.0008 N) = 2 clusters
]

MAX_KEEP = 8192


def thin(v):
    """Deterministic sub-sample of a big tensor so fixtures stay small: tensors with more
    than MAX_KEEP elements are viewed as [rows, -1] and strided (odd strides) on both axes."""
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu()
    numel = int(np.prod(v.shape)) if v.ndim else 1
    if numel <= MAX_KEEP:
        return v
    v2 = v.reshape(v.shape[0], -1) if v.ndim > 1 else v.reshape(1, -1)
    r, c = v2.shape
    cdiv = lambda a, b: (a + b - 1) // b
    sc = 1
    while r * cdiv(c, sc) > MAX_KEEP and sc < c:
        sc += 2
    sr = 1
    while cdiv(r, sr) * cdiv(c, sc) > MAX_KEEP:
        sr += 2
    return v2[::sr, ::sc]


def loss_inputs(c):
    """Positive, row-normalised maps (what softmax attention produces) and the small BatchLoss inputs; shared with the tests."""
    from dml_b200 import synth
    shp = (c["N"], 8, c["L1"], c["L2"])
    att = {k: torch.softmax(synth.normal(shp, c["seed"], k) * 2.0, dim=-1) for k in ("a1_10", "a1_20", "a2_10", "a2_20")}
    att["omic"] = synth.normal((c["N"], 128), c["seed"], "omic")
    att["vgrid"] = synth.normal((8 * c["N"], 2, 12, 12), c["seed"], "vgrid")
    return att
