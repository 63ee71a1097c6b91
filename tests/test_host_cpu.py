"""CPU: host-side logic of the drop-in surface — sampler tables, state_dict compatibility, checkpoint
converter, the C-ABI library's exports, error behaviour without a GPU, seed sharding over gloo."""
import ctypes
import os
import re
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from canon import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/sd"


# ------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from pytorch_stable_diffusion_b200 import _ext
    from pytorch_stable_diffusion_b200.csrc import build
    build.build()
    header = open(os.path.join(ROOT, "include", "sdb200.h")).read()
    declared = set(re.findall(r"\b(sdb_[a-z0-9_]+)\s*\(", header))
    declared -= {"sdb_gemm_args", "sdb_attn_args"}
    lib = ctypes.CDLL(_ext.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"not exported: {missing}"
    unbound = sorted(declared - set(_ext.SIGNATURES))
    assert not unbound, f"declared in sdb200.h but not bound in _ext.py: {unbound}"
    assert _ext.lib().sdb_abi_version() == 2
    # no torch types at the boundary: the library links only the CUDA runtime
    assert ctypes.sizeof(_ext.GemmArgs) > 0 and ctypes.sizeof(_ext.AttnArgs) > 0


def test_argument_structs_match_the_library_and_the_header_and_the_docs():
    """ctypes structs == sizeof() in the built library == field order of include/sdb200.h == INTEGRATION.md."""
    import subprocess
    from pytorch_stable_diffusion_b200 import _ext
    lib = _ext.lib()
    assert lib.sdb_args_size(0) == ctypes.sizeof(_ext.GemmArgs)
    assert lib.sdb_args_size(1) == ctypes.sizeof(_ext.AttnArgs)
    assert lib.sdb_args_size(7) == -1
    header = open(os.path.join(ROOT, "include", "sdb200.h")).read()
    for struct, c_name in ((_ext.GemmArgs, "sdb_gemm_args"), (_ext.AttnArgs, "sdb_attn_args")):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (c_name, c_name), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", first)[-1])
            names += [r.strip().lstrip("*").strip() for r in rest]
        assert names == [f[0] for f in struct._fields_], f"{c_name}: header and _ext.py disagree"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_integration.py"), "--check"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_c_abi_reports_bad_arguments_without_a_gpu():
    from pytorch_stable_diffusion_b200 import _ext
    lib = _ext.lib()
    rc = lib.sdb_layernorm(None, None, None, None, 0, 0, 0.0, 0, 0, None)
    assert rc == -1 and b"sdb_layernorm" in lib.sdb_last_error()
    with pytest.raises(ValueError):
        _ext.check(rc, "sdb_layernorm")


def test_product_path_has_no_cpu_fallback():
    from pytorch_stable_diffusion_b200 import attention, pipeline
    from pytorch_stable_diffusion_b200.synthetic import StubTokenizer
    with pytest.raises(RuntimeError):
        attention.SelfAttention(2, 64)(torch.zeros(1, 8, 64))
    with pytest.raises(RuntimeError):
        pipeline.generate("a", "b", models={}, tokenizer=StubTokenizer(), device="cpu")
    with pytest.raises(ValueError):
        pipeline.generate("a", "b", strength=1.5, models={}, tokenizer=StubTokenizer(), device="cpu")
    with pytest.raises(ValueError):
        pipeline.generate("a", "b", sampler_name="ddim", models={}, tokenizer=StubTokenizer(), device="cpu")
    with pytest.raises(ValueError, match="largest supported image"):
        pipeline.generate("a", "b", models={}, tokenizer=StubTokenizer(), device="cuda", height=2048, width=2048)


def test_unsafe_pickle_is_opt_in(tmp_path, monkeypatch):
    """A checkpoint that needs arbitrary unpickling is refused unless the caller opts in (the reference's
    weights_only=False, sd/model_converter.py:5, executes whatever the file contains)."""
    import pickle
    from pytorch_stable_diffusion_b200 import model_converter

    import fractions                   # any class outside torch's allow-list: weights_only=True rejects it
    path = str(tmp_path / "legacy.ckpt")
    torch.save({"state_dict": {}, "extra": fractions.Fraction(1, 3)}, path)
    monkeypatch.delenv("SDB_ALLOW_UNSAFE_PICKLE", raising=False)
    with pytest.raises(pickle.UnpicklingError, match="allow_unsafe_pickle"):
        model_converter.load_from_standard_weights(path, "cpu")
    with pytest.raises(KeyError):      # opted in: the file is read (and then found to hold no SD weights)
        model_converter.load_from_standard_weights(path, "cpu", allow_unsafe_pickle=True)


def test_graph_cache_is_a_small_lru():
    from pytorch_stable_diffusion_b200 import pipeline
    assert 2 <= pipeline.GRAPH_CACHE_SIZE <= 4
    assert isinstance(pipeline._GRAPH_CACHE, __import__("collections").OrderedDict)


# ------------------------------------------------------------------------------------ sampler
def test_ddpm_sampler_matches_reference_golden():
    from pytorch_stable_diffusion_b200.ddpm import DDPMSampler
    d = golden("blocks.pt")["ddpm"]
    s = DDPMSampler(torch.Generator().manual_seed(5))
    assert abs(float(s.betas[0]) - 8.5e-5) < 1e-9          # the reference's beta_start default
    s.set_inference_timesteps(50)
    assert torch.equal(s.timesteps, d["timesteps"]) and s.timesteps.dtype == torch.int64
    assert torch.equal(s.alphas_cumprod, d["alphas_cumprod"])
    assert torch.equal(s._get_variance(980), d["var_980"]) and torch.equal(s._get_variance(0), d["var_0"])
    lat, mo = d["step_in"]
    assert torch.allclose(s.step(980, lat, mo), d["step_980"], rtol=0, atol=1e-6)
    assert torch.allclose(s.step(0, lat, mo), d["step_0"], rtol=0, atol=1e-6)
    table = s.coefficient_table()
    assert table.shape == (50, 5) and float(table[-1, 4]) == 0.0
    s.set_strength(0.8)
    assert s.start_step == d["strength08"][0] and torch.equal(s.timesteps, d["strength08"][1])
    assert s.coefficient_table().shape == (40, 5)


def test_add_noise_matches_formula():
    from pytorch_stable_diffusion_b200.ddpm import DDPMSampler
    s = DDPMSampler(torch.Generator().manual_seed(9))
    s.set_inference_timesteps(50)
    x = torch.randn(2, 4, 8, 8)
    out = s.add_noise(x, torch.tensor([780]))
    z = torch.randn(x.shape, generator=torch.Generator().manual_seed(9))
    a = s.alphas_cumprod[780]
    assert torch.allclose(out, a ** 0.5 * x + (1 - a) ** 0.5 * z, atol=1e-6)


# ------------------------------------------------------------------------------------ state dicts
def test_state_dict_keys_match_the_reference_surface():
    from pytorch_stable_diffusion_b200 import clip, decoder, diffusion, encoder
    with torch.device("meta"):
        mods = {"diffusion": diffusion.Diffusion(), "decoder": decoder.VAE_Decoder(),
                "encoder": encoder.VAE_Encoder(), "clip": clip.CLIP()}
    counts = {k: len(m.state_dict()) for k, m in mods.items()}
    assert counts == {"diffusion": 654, "decoder": 136, "encoder": 104, "clip": 148}
    sd = mods["diffusion"].state_dict()
    assert sd["unet.encoders.1.1.attention_1.in_proj.weight"].shape == (960, 320)
    assert "unet.encoders.1.1.attention_1.in_proj.bias" not in sd
    assert sd["unet.encoders.1.1.linear_geglu_1.weight"].shape == (2560, 320)
    assert sd["unet.decoders.2.1.conv.weight"].shape == (1280, 1280, 3, 3)
    assert "3.groupnorm.weight" in mods["decoder"].state_dict()       # declared, never applied
    assert mods["clip"].state_dict()["embedding.position_embedding"].shape == (77, 768)
    g = golden("blocks.pt")
    diffusion.UNET_ResidualBlock(64, 128).load_state_dict(g["unet_res"]["sd"], strict=True)
    diffusion.UNET_AttentionBlock(2, 32).load_state_dict(g["unet_attn"]["sd"], strict=True)
    decoder.VAE_ResidualBlock(64, 128).load_state_dict(g["vae_res"]["sd"], strict=True)
    decoder.VAE_AttentionBlock(64).load_state_dict(g["vae_attn"]["sd"], strict=True)
    clip.CLIPLayer(4, 64).load_state_dict(g["clip_layer"]["sd"], strict=True)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_state_dict_keys_and_shapes_are_identical():
    import importlib
    from pytorch_stable_diffusion_b200 import clip, decoder, diffusion, encoder
    sys.path.insert(0, REF)
    try:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules)
                 if k in ("attention", "clip", "decoder", "diffusion", "encoder", "ddpm", "pipeline")}
        ref = {n: importlib.import_module(n) for n in ("clip", "decoder", "diffusion", "encoder")}
        with torch.device("meta"):
            pairs = [(diffusion.Diffusion(), ref["diffusion"].Diffusion()),
                     (decoder.VAE_Decoder(), ref["decoder"].VAE_Decoder()),
                     (encoder.VAE_Encoder(), ref["encoder"].VAE_Encoder()),
                     (clip.CLIP(), ref["clip"].CLIP())]
        for mine, theirs in pairs:
            a, b = mine.state_dict(), theirs.state_dict()
            assert list(a) == list(b)
            assert all(a[k].shape == b[k].shape for k in a)
    finally:
        sys.path.remove(REF)
        for n in ("attention", "clip", "decoder", "diffusion", "encoder", "ddpm", "pipeline"):
            sys.modules.pop(n, None)
        sys.modules.update(saved)


# ------------------------------------------------------------------------------------ converter
def _synthetic_checkpoint():
    """Tiny tensors under every CompVis key the converter reads (shapes only matter for the fused
    q/k/v concatenations and the two VAE reshapes)."""
    from pytorch_stable_diffusion_b200 import model_converter
    sd = {}
    g = torch.Generator().manual_seed(0)
    for rules in model_converter.conversion_rules().values():
        for dst, (op, srcs) in rules:
            for s in srcs:
                if s in sd:
                    continue
                if op in ("cat_matrix", "matrix"):     # the reference reshapes these to (1536|512, 512)
                    sd[s] = torch.randn(512, 512, 1, 1, generator=g)
                elif s.endswith(".weight"):
                    sd[s] = torch.randn(6, 3, generator=g)
                else:
                    sd[s] = torch.randn(6, generator=g)
    return sd


def test_converter_covers_every_destination_key():
    from pytorch_stable_diffusion_b200 import clip, decoder, diffusion, encoder, model_converter
    out = model_converter.convert_state_dict(_synthetic_checkpoint())
    with torch.device("meta"):
        mods = {"diffusion": diffusion.Diffusion(), "decoder": decoder.VAE_Decoder(),
                "encoder": encoder.VAE_Encoder(), "clip": clip.CLIP()}
    for k, m in mods.items():
        assert set(out[k]) == set(m.state_dict()), k
    assert out["decoder"]["3.attention.in_proj.weight"].shape == (1536, 512)
    assert out["decoder"]["3.attention.out_proj.weight"].shape == (512, 512)
    assert out["diffusion"]["unet.encoders.1.1.attention_1.in_proj.weight"].shape == (18, 3)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_converter_equals_reference_converter(tmp_path):
    import importlib.util
    from pytorch_stable_diffusion_b200 import model_converter
    ckpt = _synthetic_checkpoint()
    path = str(tmp_path / "synthetic.ckpt")
    torch.save({"state_dict": ckpt}, path)
    spec = importlib.util.spec_from_file_location("ref_model_converter", os.path.join(REF, "model_converter.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    theirs = ref.load_from_standard_weights(path, "cpu")
    mine = model_converter.load_from_standard_weights(path, "cpu")
    assert set(mine) == set(theirs)
    for grp in theirs:
        assert set(mine[grp]) == set(theirs[grp]), grp
        for k, v in theirs[grp].items():
            assert torch.equal(mine[grp][k], v), (grp, k)


# ------------------------------------------------------------------------------------ sharding (gloo)
def test_shard_range_partitions_seeds():
    from pytorch_stable_diffusion_b200.sharding import shard_range, shard_seeds
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_seeds(list(range(42, 106)), 3, 8) == list(range(66, 74))     # config 4: 64 seeds / 8 GPUs
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from pytorch_stable_diffusion_b200.sharding import gather_images, shard_seeds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    seeds = list(range(100, 100 + n_total))
    mine = shard_seeds(seeds)
    # stand-in for the sampler: image i is filled with its seed (the GPU path is covered by -m gpu)
    local = torch.stack([torch.full((4, 4, 3), s % 256, dtype=torch.uint8) for s in mine]) if mine else \
        torch.zeros((0, 4, 4, 3), dtype=torch.uint8)
    allimg = gather_images(local, n_total)
    q.put((rank, mine, allimg[:, 0, 0, 0].tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 5])
def test_seed_sharding_and_gather_world_size_2_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seeds = list(range(100, 100 + n_total))
    assert res[0][1] + res[1][1] == seeds                       # disjoint, ordered, complete
    for _, _, gathered in res:
        assert gathered == [s % 256 for s in seeds]             # every rank sees all images in seed order


# ------------------------------------------------------------------------------------ pack-time algebra (CPU)
def test_upsample_phase_weights_reproduce_upsample_then_conv():
    """engine.pack_upsample_phases: conv3x3(nearest_upsample_x2(x)) == four parity-phase 2x2 convolutions of x with the
    3x3 taps that land on the same input pixel summed (sd/diffusion.py:412-435) - checked in fp32 on the CPU."""
    import torch.nn.functional as F
    from pytorch_stable_diffusion_b200 import engine
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 6, 3, padding=1)
    x = torch.randn(2, 8, 5, 7)
    with torch.no_grad():
        ref = conv(F.interpolate(x, scale_factor=2, mode="nearest"))
        w4, b = engine.pack_upsample_phases(conv, "cpu", torch.float32)
        out = torch.zeros_like(ref)
        xp = F.pad(x, (1, 1, 1, 1))
        for a in (0, 1):
            for bb in (0, 1):
                w = w4[2 * a + bb].view(6, 2, 2, 8).permute(0, 3, 1, 2)          # [Cout, Cin, u, v]
                out[:, :, a::2, bb::2] = F.conv2d(xp[:, :, a:a + 6, bb:bb + 8], w) + b.view(1, -1, 1, 1)
    assert float((out - ref).abs().max()) < 1e-5


def test_resample_tables_reproduce_pillow():
    """imageio.resample_tables (the taps sdb_resample_u8 runs on the device) evaluated in numpy against
    PIL.Image.resize on the reference's test image: byte-exact for up- and down-sampling."""
    import numpy as np
    from PIL import Image
    from pytorch_stable_diffusion_b200 import imageio
    g = golden("img2img_5.pt")
    if g is None:
        pytest.skip("img2img golden (holds dog.jpg) not generated")
    img = g["input"].numpy()
    dog = Image.fromarray(img)

    def one_pass(a, out_size, axis):
        bounds, coef = imageio.resample_tables(a.shape[axis], out_size)
        src = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.zeros((out_size,) + src.shape[1:], np.int64)
        for o in range(out_size):
            lo, cnt = bounds[o]
            acc = np.full(src.shape[1:], 1 << 21, np.int64)
            for k in range(cnt):
                acc += src[lo + k] * int(coef[o, k])
            out[o] = np.clip(acc >> 22, 0, 255)
        return np.moveaxis(out, 0, axis).astype(np.uint8)

    for w, h in ((768, 768), (300, 200), (1024, 520)):
        got = one_pass(one_pass(img, w, 1), h, 0)
        assert np.array_equal(got, np.array(dog.resize((w, h)))), (w, h)


def test_tile_chooser_rules():
    """ops._choose_tiling is pure host arithmetic: the rules the GPU measurements fixed hold -
    160-column tiles for the UNet's channel counts, 256-column tiles for 16-bit-only results from 640 channels up (the
    TMA epilogue then stores 64-column units), split-K only for long reductions on few row tiles, never a tile wider
    than the output, and every choice is a width the kernel accepts."""
    from pytorch_stable_diffusion_b200 import ops
    valid = lambda bn: (bn % 16 == 0 and 16 <= bn <= 256) or bn == 320
    # 16-bit result only, short reduction: 64-column store units want tile widths that are multiples of 64
    assert ops._choose_tiling(65536, 1280, 5, 2, 0) == (256, 1)
    assert ops._choose_tiling(16384, 640, 10, 2, 2) == (256, 1)
    assert ops._choose_tiling(65536, 320, 5, 2, 2) == (160, 1)
    # an fp32 tensor or an fp32 residual is involved: the 160-column tiles of the timelines
    assert ops._choose_tiling(65536, 640, 10, 4 + 2, 4)[0] == 160
    assert ops._choose_tiling(65536, 320, 45, 4, 0) == (160, 1)
    # the 8 x 8 level: 4 row-tile pairs for 74 slots -> the long reduction is split
    bn, ns = ops._choose_tiling(1024, 1280, 180, 4, 0)
    assert ns > 1 and valid(bn)
    for rows in (77, 1024, 4096, 16384, 65536, 147456):
        for cout in (4, 48, 96, 320, 640, 768, 1280, 2560):
            for nkb in (1, 5, 20, 45, 180):
                for ob, rb in ((2, 0), (2, 2), (4, 4), (6, 4)):
                    bn, ns = ops._choose_tiling(rows, cout, nkb, ob, rb)
                    assert valid(bn) and ns >= 1 and ns <= max(1, nkb), (rows, cout, nkb, ob, rb, bn, ns)
                    assert bn <= max(16, (cout + 15) // 16 * 16) or bn == 320 and cout % 320 == 0


def test_header_is_plain_c_and_matches_the_bindings(tmp_path):
    """include/sdb200.h is the boundary a non-Python host binds to: it must compile as C99 on its own (no C++ / CUDA /
    torch types), warning-free, and the struct sizes a C compiler computes must equal the ctypes mirrors in _ext.py."""
    import ctypes
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from pytorch_stable_diffusion_b200 import _ext
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "sdb200.h"\n'
                   'int main(void) { printf("%d %zu %zu\\n", SDB_ABI_VERSION, sizeof(struct sdb_gemm_args), '
                   'sizeof(struct sdb_attn_args)); return 0; }\n')
    exe = tmp_path / "abi"
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ver, gemm, attn = (int(t) for t in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    assert gemm == ctypes.sizeof(_ext.GemmArgs) and attn == ctypes.sizeof(_ext.AttnArgs)
    assert ver == _ext.lib().sdb_abi_version()        # the library was built from this header


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h"):
        pytest.skip("no gcc / CUDA headers")
    from pytorch_stable_diffusion_b200 import _ext
    _ext.lib()                                   # raises if the library is not built
    csrc = os.path.join(ROOT, "pytorch_stable_diffusion_b200", "csrc")
    exe = tmp_path / "abi_linear"
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-I", "/usr/local/cuda/include", os.path.join(ROOT, "examples", "abi_linear.c"), "-L", csrc,
                        "-lsdb200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{csrc}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_program_builds_against_the_library(tmp_path):
    """examples/abi_linear.c - a C99 program with no Python and no torch - compiles and links against include/sdb200.h,
    libsdb200.so and the CUDA runtime alone (it runs in tests/test_kernels_gpu.py::test_c_host_program)."""
    _build_c_example(tmp_path)
