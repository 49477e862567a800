"""GPU: run the reference's OWN backend-vs-CPU parity harness (ggml/tests/test-backend-ops.cpp, compiled from
/root/reference into oracle/_ref/) against libggml-b200.so, loaded exactly as a user would: GGML_BACKEND_PATH."""
import os
import re
import subprocess

import pytest

from oracle.cpu_ref import test_backend_ops as tbo_path
from sdb200 import B200_SO

pytestmark = pytest.mark.gpu

GROUPS = {
    "elementwise": "ADD,SUB,MUL,DIV,SCALE,CLAMP,SQR,SQRT,SIN,COS,LOG,LEAKY_RELU,SILU,GELU,GELU_QUICK,RELU,SIGMOID,TANH,EXP,NEG,ABS,GELU_ERF,HARDSWISH,HARDSIGMOID,STEP,SGN,ELU",
    "movement": "CPY,CONT,DUP,CONCAT,REPEAT,PAD,UPSCALE,TIMESTEP_EMBEDDING,GET_ROWS,ARANGE,SUM_ROWS,MEAN,IM2COL,IM2COL_3D",
    "norms": "GROUP_NORM,NORM,RMS_NORM,L2_NORM,SOFT_MAX",
    "mul_mat": "MUL_MAT",
    "flash_attn": "FLASH_ATTN_EXT",
}


@pytest.mark.parametrize("group", list(GROUPS))
def test_reference_backend_ops_harness(group):
    env = dict(os.environ, GGML_BACKEND_PATH=str(B200_SO))
    r = subprocess.run([str(tbo_path()), "test", "-b", "B200_0", "-o", GROUPS[group]], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=1500)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
    fails = [l for l in out.splitlines() if "FAIL" in l]
    m = re.search(r"(\d+)/(\d+) tests passed", out)
    assert m, out[-2000:]
    assert int(m.group(2)) > 0, "no test case ran: device not found or every case unsupported\n" + out[-1500:]
    assert m.group(1) == m.group(2) and r.returncode == 0, f"{m.group(0)}\n" + "\n".join(fails[:40])
