"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports every declared symbol."""
import ctypes
import importlib
import os

import pytest
import torch


def test_library_builds_and_exports_every_declared_symbol():
    _lib = importlib.import_module("sln_b200._lib")
    path = _lib.build()
    lib = ctypes.CDLL(path)
    declared = _lib.declared_symbols()
    assert len(declared) >= 15
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, "symbols declared in include/sln_b200.h but not exported: %s" % missing
    unbound = [s for s in declared if s not in _lib.SIGNATURES]
    assert not unbound, "symbols without a ctypes signature: %s" % unbound


def test_version_and_error_string():
    _lib = importlib.import_module("sln_b200._lib")
    lib = _lib.load()
    assert lib.sln_version() == 2
    # argument validation happens on the host before any launch: no GPU needed
    d = _lib.VaeDesc(embedding_dim=6, n_layers=5, recurrent=0, norm=1, training=1, box_dim=6, n_angle=24, num_objs=33, num_preds=16,
                     num_attrs=5, bn_eps=1e-5, bn_momentum=0.1, gconv_dim_override=0, gconv_hidden_override=0)
    assert lib.sln_vae_num_params(d) == -1
    assert b"multiple of 4" in lib.sln_last_error()


def test_parameter_table_matches_module_tree():
    from helpers import our_model
    _lib = importlib.import_module("sln_b200._lib")
    lib = _lib.load()
    for norm, mode, layers in (("batch", "feedforward", 5), ("none", "feedforward", 5), ("batch", "recurrent", 3)):
        m = our_model(E=64, layers=layers, norm=norm, mode=mode)
        ps, bufs, enc, dec = m._param_list()
        assert lib.sln_vae_num_params(m._desc()) == len(ps)
        assert lib.sln_vae_num_bn(m._desc()) * 3 == len(bufs)
        assert sorted(enc + dec) == list(range(len(ps)))
        assert {id(p) for p in ps} == {id(p) for p in m.parameters()}
        assert lib.sln_vae_workspace_bytes(m._desc(), 2048, 3968, 0) > 0


def test_cpu_tensors_fail_loudly():
    from helpers import our_model, syn
    m = our_model(E=8, layers=1)
    objs, triples, boxes, angles, attrs = syn.fixture_graph()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(objs, triples, boxes, angles, attrs, None)
    graph = importlib.import_module("sln_b200.models.graph")
    g = graph.GraphTripleConv(16, hidden_dim=32)
    with pytest.raises(RuntimeError, match="CUDA"):
        g(torch.zeros(4, 16), torch.zeros(3, 16), torch.zeros(3, 2, dtype=torch.long))


def test_state_dict_keys_and_init_match_reference_when_available():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    from helpers import our_model, syn
    Ref = ref_shim.vae_model_class()
    for norm in ("batch", "none"):
        torch.manual_seed(42)
        r = Ref(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
                gconv_num_layers=5, mlp_normalization=norm, vec_noise_dim=0, layout_noise_dim=32, use_AE=False)
        o = our_model(E=64, layers=5, norm=norm)
        a, b = r.state_dict(), o.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_host_side_argument_checks_of_collate_refine_scene_entry_points():
    """Argument validation happens on the host before any launch, so the error convention (negative code + message, no exception
    across the ABI) can be checked without a GPU: collate wire layout, refinement-loss pyramid limits, scene assembly / compositing."""
    lib = importlib.import_module("sln_b200._lib").load()
    off = (ctypes.c_int64 * 10)()
    assert lib.sln_collate_layout(3, 50, 75, 6, off) == 0
    o = list(off)
    assert o == sorted(o) and all(v % 16 == 0 for v in o)                       # sections ascending, 16-byte aligned
    assert o[1] - o[0] >= 8 * 3 and o[5] - o[4] >= 8 * 50 and o[8] - o[7] >= 8 * 3 * 75 and o[9] - o[8] >= 4 * 6 * 50
    assert lib.sln_collate_layout(-1, 0, 0, 6, off) < 0 and b"collate_layout" in lib.sln_last_error()
    assert lib.sln_collate_finish(None, 0, 1, 1, 1, 6, None, None, None, None, None) < 0
    sizes = (ctypes.c_int32 * 4)(32, 48, 64, 96)
    assert lib.sln_refine_loss_workspace_bytes(256, sizes, 40, 29) > 4 * 69 * (32 * 32 + 48 * 48 + 64 * 64 + 96 * 96)
    assert lib.sln_refine_loss_workspace_bytes(256, sizes, 0, 29) == 0
    dummy = ctypes.c_void_p(256)                                                 # never dereferenced: the checks fail first
    counts = (ctypes.c_float * 4)(1, 1, 1, 1)
    too_coarse = (ctypes.c_int32 * 4)(8, 48, 64, 96)                             # 96 / 8 = 12 destinations per source: beyond the tap lists
    assert lib.sln_refine_loss(dummy, 256, too_coarse, 40, 29, dummy, dummy, counts, dummy, None, dummy, 1 << 30, None) < 0
    assert b"taps" in lib.sln_last_error()
    assert lib.sln_refine_loss(dummy, 256, sizes, 65, 29, dummy, dummy, counts, dummy, None, dummy, 1 << 30, None) < 0   # n_sem > 64
    assert lib.sln_refine_loss(dummy, 256, sizes, 40, 29, dummy, dummy, counts, dummy, None, dummy, 16, None) == -2      # SLN_EWORKSPACE
    assert lib.sln_scene_assemble_workspace_bytes(10) >= 10 * 32
    assert lib.sln_scene_assemble_fwd(None, None, 1, None, 0, None, None, None, 0, None, 0, None, None, None, 0, None, None, 0.06, None, None,
                                      None, None, 0, None) < 0
    assert lib.sln_composite_workspace_bytes(32) > 0
    assert lib.sln_composite_fwd(dummy, dummy, 32, 65536, 40, dummy, 41, dummy, 29, dummy, dummy, dummy, 1 << 20, None) < 0   # wall >= C
    assert b"composite_fwd" in lib.sln_last_error()


def test_ctypes_struct_mirrors_match_the_header_layout(tmp_path):
    """sln_vae_desc / sln_bn_sync are passed by pointer across the C ABI: the ctypes mirrors in _lib.py must have the size and the field
    offsets that a C compiler gives the header's structs (a field added on one side only would shift every later field silently)."""
    import ctypes
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    _lib = importlib.import_module("sln_b200._lib")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mirrors = {"sln_vae_desc": _lib.VaeDesc, "sln_bn_sync": _lib.BnSync}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sln_b200.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)   # the header is plain C
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        if not line:
            continue
        cname, fname, val = line.split()
        cls = mirrors[cname]
        if fname == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, fname).offset == int(val), (cname, fname, getattr(cls, fname).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in mirrors.values())
