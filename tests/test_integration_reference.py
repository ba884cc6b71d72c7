"""Drop-in boundary against the REAL reference (only where /root/reference exists, i.e. the build container; skipped
on the GPU box): shadowing the module names as INTEGRATION.md section 1 describes, the reference's own unmodified
models/model.py must build DeformPathomicNet on top of the dml_b200 operators, with exactly the reference's
state_dict keys and shapes."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, types, json
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
    import torch
    from oracle.make_goldens import install_reference_shims      # stubs only ABSENT third-party imports
    install_reference_shims()
    from types import SimpleNamespace
    args = SimpleNamespace(path_dim=128, omic_dim=128, mmhid=128, attn_dim=1, return_vgrid=False, label_dim=4,
                           input_size_omic_tumor=59, input_size_omic_immune=361, return_grad="False", dropout_rate=0.1,
                           init_type="max", fusion_type="concat", task_type="diag2021", mode="deformpathomic",
                           input_size_omic=431, act_type="none", use_bilinear=1, skip=1, gpu_ids="0", use_sparsemax=0,
                           init_gain=0.02)
    def build(shadow):
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        if shadow:                                                  # INTEGRATION.md section 1
            from dml_b200 import DeformableAttention1D, DeformCrossTransMIL, NystromAttention as _nys, mil
            sys.modules["models.DeformableAttention1D"] = DeformableAttention1D
            sys.modules["models.DeformCrossTransMIL"] = DeformCrossTransMIL
            sys.modules["models.mil"] = mil
            sys.modules["nystrom_attention"] = types.SimpleNamespace(NystromAttention=_nys.NystromAttention)
        from models.model import define_net
        net = define_net(args)
        mods = {type(m).__module__ for m in net.modules()}
        return {k: list(v.shape) for k, v in net.state_dict().items()}, sorted(mods)
    ref_sd, ref_mods = build(False)
    our_sd, our_mods = build(True)
    print(json.dumps({"same": ref_sd == our_sd, "n": len(our_sd), "ours_used": any(m.startswith("dml_b200") for m in our_mods),
                      "ref_clean": not any(m.startswith("dml_b200") for m in ref_mods)}))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_reference_model_py_builds_on_shadowed_operators():
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "ref": REF}], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["same"] and res["ours_used"] and res["ref_clean"] and res["n"] == 118, res


SCRIPT_COATTN = textwrap.dedent('''
    import sys, types, json
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
    import torch
    from oracle.make_goldens import install_reference_shims
    install_reference_shims()
    from types import SimpleNamespace
    args = SimpleNamespace(path_dim=128, omic_dim=128, mmhid=128, attn_dim=1, return_vgrid=False, label_dim=4,
                           input_size_omic_tumor=59, input_size_omic_immune=361, return_grad="False", dropout_rate=0.1,
                           init_type="max", fusion_type="concat", task_type="survival", mode="mcat",
                           input_size_omic=431, act_type="none", use_bilinear=1, skip=1, gpu_ids="0", use_sparsemax=0,
                           init_gain=0.02)
    def build(shadow, mode):
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils.loss"]:
            del sys.modules[k]
        if shadow:                                                  # INTEGRATION.md section 1
            from dml_b200 import MultiheadAttention as _mha, loss as _loss
            sys.modules["models.MultiheadAttention"] = _mha
            sys.modules["utils.loss"] = _loss
        from models.model import define_net
        from utils.loss import BatchLoss, PathBatchLoss, OmicDomainScaleLoss, DistillationLoss      # train_test.py:9
        args.mode = mode
        net = define_net(args)
        mods = {type(m).__module__ for m in net.modules()}
        return {k: list(v.shape) for k, v in net.state_dict().items()}, sorted(mods), PathBatchLoss.__module__
    res = {}
    for mode in ("mcat", "cmta"):
        ref_sd, ref_mods, _ = build(False, mode)
        our_sd, our_mods, lm = build(True, mode)
        res[mode] = {"same": ref_sd == our_sd, "n": len(our_sd), "ours_used": "dml_b200.MultiheadAttention" in our_mods,
                     "ref_clean": not any(m.startswith("dml_b200") for m in ref_mods), "loss_module": lm}
    print(json.dumps(res))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_reference_mcat_and_cmta_build_on_the_shadowed_coattention():
    """the unmodified models/model.py builds MCAT_Surv and CMTA on dml_b200.MultiheadAttention (same state_dict
    keys and shapes) and train_test.py's `from utils.loss import ...` resolves to the dml_b200 mirror."""
    out = subprocess.run([sys.executable, "-c", SCRIPT_COATTN % {"root": ROOT, "ref": REF}], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads(out.stdout.strip().splitlines()[-1])
    for mode, n in (("mcat", 92), ("cmta", 104)):
        r = res[mode]
        assert r["same"] and r["ours_used"] and r["ref_clean"] and r["n"] == n and r["loss_module"].startswith("dml_b200"), (mode, r)


SCRIPT_TEACHER = textwrap.dedent('''
    import sys, types, json
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
    import torch
    from oracle.make_goldens import install_reference_shims
    install_reference_shims()
    from types import SimpleNamespace
    args = SimpleNamespace(path_dim=128, omic_dim=128, mmhid=128, attn_dim=2, return_vgrid=False, label_dim=4,
                           input_size_omic_tumor=59, input_size_omic_immune=361, return_grad="False", dropout_rate=0.1,
                           init_type="max", fusion_type="concat", task_type="survival", mode="teacher",
                           input_size_omic=431, act_type="none", use_bilinear=1, skip=1, gpu_ids="0", use_sparsemax=0,
                           init_gain=0.02, path_cluster_num=0.0008, omic_cluster_num=2, combination_type="max_confidence",
                           combination_type_teas="max_confidence", combination_type_stus="max_confidence", path_scale=1, path_gate=1,
                           omic_scale=1, omic_gate=1, cut_fuse_grad=False)
    def build(shadow, mode):
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        if shadow:                                                  # INTEGRATION.md section 1
            from dml_b200 import DeformableAttention2D as _d2, ClusterMergeNet as _cm
            sys.modules["models.DeformableAttention2D"] = _d2
            sys.modules["models.ClusterMergeNet"] = _cm
        from models.model import define_net
        args.mode = mode
        net = define_net(args)
        mods = {type(m).__module__ for m in net.modules()}
        return {k: list(v.shape) for k, v in net.state_dict().items()}, sorted(mods)
    res = {}
    for mode in ("teacher", "student"):
        ref_sd, ref_mods = build(False, mode)
        our_sd, our_mods = build(True, mode)
        res[mode] = {"same": ref_sd == our_sd, "n": len(our_sd), "ours": [m for m in our_mods if m.startswith("dml_b200")],
                     "ref_clean": not any(m.startswith("dml_b200") for m in ref_mods)}
    print(json.dumps(res))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_reference_teacher_and_student_build_on_the_shadowed_2d_attention():
    """the unmodified models/model.py + models/Modules.py build the teacher and student nets (the variant the
    shipped YAMLs select: attn_dim 2) on dml_b200.DeformableAttention2D / ClusterMergeNet with identical state_dict keys and shapes."""
    out = subprocess.run([sys.executable, "-c", SCRIPT_TEACHER % {"root": ROOT, "ref": REF}], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert "dml_b200.DeformableAttention2D" in res["teacher"]["ours"], res
    assert "dml_b200.ClusterMergeNet" in res["student"]["ours"] and "dml_b200.DeformableAttention2D" in res["student"]["ours"], res
    for mode in ("teacher", "student"):
        r = res[mode]
        assert r["same"] and r["ref_clean"] and r["n"] > 50, (mode, r)
