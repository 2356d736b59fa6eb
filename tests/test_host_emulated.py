"""CPU suite: the product's HOST logic (everything above the C ABI) against the fp32 oracle, with the CUDA kernels replaced by
their torch restatements (tests/emu_ops.py).  The same checkers run against the real kernels in the `-m gpu` suite; here they
prove that the tape engine, the LoRA slot packing, the v1 / V2 / post_add / concat_hidden control algebra, the hint-encoder
program and the Trainer issue the right operations - on a box without a GPU.  One subprocess runs all cases (the emulation mode
is a per-process switch, tests/_device.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CASES = ["unet_none", "unet_plain", "unet_v1", "unet_v2", "unet_v1_stacked", "unet_v1_post_add", "unet_v1_concat",
         "unet_v1_stacked@0.5", "unet_v2@0.5", "hint_v1", "hint_v2", "train_v1", "train_v2", "train_lora_only", "resume", "accumulate", "vae_tiny", "clip_tiny"]
# general adapter chains (tests/check_variants.py) and the diffusers-style per-module processor call (tests/check_eager.py); listed
# by name so that collecting this file never imports a checker (the device mode is fixed at their import, tests/_device.py)
CASES += ["variant_" + k for k in ("v1_pre_post_add", "v1_post_add_main_stacked", "v2_post_post_add", "v1_rank16", "v1_rank8_stacked8",
                                   "v1_control_rank12", "v2_control_rank8", "post_add_rank8", "v1_on_v1", "v2_on_v2",
                                   "v1_concat_stacked", "v1_concat_post_add", "v1_concat_rank8")]
CASES += ["cfg_broadcast_v1", "cfg_broadcast_v2"]                                      # control batch 1 under a CFG UNet batch of 2
CASES += ["two_forwards_v1_stacked", "two_forwards_v2", "unet_v1_stacked@0.0"]      # two UNet calls (different scale) before one backward; scale 0
CASES += ["generic_" + v for v in ("plain", "v1", "v2", "v1_stacked@0.5", "v1_post_add", "v1_concat")]
CASES += ["refgold_" + k for k in ("v1_stacked", "v2", "post_add", "concat")]      # vs vectors computed by the reference's own models.py
CASES += ["sampler_ddim", "sampler_dpmpp", "step_glue", "step_from_pixels", "generate", "graphed_program"]                              # tests/check_sampler.py
CASES += ["eager_" + k for k in ("plain_self", "plain_cross", "v1_self", "v1_cross_stacked", "v2_self", "v2_cross", "lora_linear")]


@pytest.fixture(scope="module")
def emulated_results():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "run_emulated.py"), *CASES], capture_output=True, text=True, timeout=1500,
                       cwd=str(ROOT))
    res = {}
    for line in r.stdout.splitlines():
        if line.startswith("RESULT "):
            _, name, verdict, _ = line.split()
            res[name] = verdict == "OK"
    return res, r.stdout[-6000:] + r.stderr[-3000:]


@pytest.mark.parametrize("case", CASES)
def test_host_logic_matches_oracle(emulated_results, case):
    res, log = emulated_results
    assert case in res, f"{case} did not run:\n{log}"
    assert res[case], f"{case} deviates from the oracle:\n{log}"


def test_emulation_is_test_only():
    """The torch restatements are test infrastructure: nothing that ships may import them."""
    for p in [*ROOT.glob("controllora_b200/*.py"), ROOT / "bench.py"]:
        src = p.read_text()
        assert "emu_ops" not in src and "run_emulated" not in src, p


def test_fullsize_sd15_shapes_pass_the_real_launchers_argument_validation():
    """One train step at the BASELINE shapes per path: canny-v2 on the fused path (the bench configuration) and v1 + a stacked
    pre-LoRA forced through the general chain path - every launch's arguments are checked by the real C launchers
    (tests/run_fullsize_emulated.py)."""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "run_fullsize_emulated.py"), "diffusiondb-canny-v2", "0", "0",
                        "diffusiondb-canny", "1", "1"], capture_output=True, text=True, timeout=1500, cwd=str(ROOT))
    lines = [l for l in r.stdout.splitlines() if l.startswith("FULLSIZE")]
    assert r.returncode == 0 and len(lines) == 2 and all(l.endswith("OK") for l in lines), r.stdout[-3000:] + r.stderr[-3000:]
    assert "'v2': 32" in lines[0] and "'generic': 32" in lines[1]
