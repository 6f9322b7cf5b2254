"""The C++ multi-GPU layer (csrc/sccg_mgpu.cuh) on real devices (`-m gpu`).
  * one GPU: the ranks are host threads with one context each on device 0 and the in-process hub as transport
    (SCCG_MGPU_HUB=1; NCCL cannot place two ranks on one device) -- the product kernels and the host logic;
  * two or more GPUs (skipped otherwise): one rank per device over NCCL (ncclCommInitRank / ncclAllGather / grouped
    ncclSend + ncclRecv), incl. a chromosome-sized pair sharded by segment range."""
import ctypes as C

import pytest

import oracle_lib as ol
import sccg_b200
import test_emu_mgpu as T
from sccg_genome_compression_b200 import synth

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(params=["hub-1gpu", "nccl"])
def backend(request, monkeypatch):
    if request.param == "hub-1gpu":
        monkeypatch.setenv("SCCG_MGPU_HUB", "1")
        monkeypatch.setitem(T.BACKEND, "devices", None)
    else:
        if _ngpu() < 2:
            pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
        monkeypatch.delenv("SCCG_MGPU_HUB", raising=False)
        monkeypatch.setitem(T.BACKEND, "devices", list(range(_ngpu())))
    monkeypatch.setitem(T.BACKEND, "lib", None)             # the product library
    return request.param


def _world(backend, want):
    return want if backend == "hub-1gpu" else min(want, _ngpu())


def test_assign(backend):
    T.test_assign_matches_python_lpt()


@pytest.mark.parametrize("world", [2, 3])
def test_whole_genome_gather(backend, world):
    T.test_whole_genome_gather(_world(backend, world))


@pytest.mark.parametrize("world", [2, 3])
def test_compress_sharded(backend, world):
    w = _world(backend, world)
    T.test_compress_sharded_local_identical_to_unsharded(w)
    T.test_compress_sharded_runs_across_borders(w)


def test_compress_sharded_fallbacks(backend):
    T.test_compress_sharded_fallbacks()                     # (2 ranks)


def test_decompress_sharded(backend):
    T.test_decompress_sharded_pieces(_world(backend, 3))


def test_compress_sharded_chromosome_sized(backend):
    """a 60 Mbp local pair over the ranks == the unsharded file (oracle), and the image decodes to the oracle's text"""
    n = 60_000_000
    ref, tgt = synth.local_pair(n, synth.seed_for(2, 93))
    rb, tb = ref.tobytes(), tgt.tobytes()
    rc, exp, mode = ol.orc_compress(rb, tb, b">sharded chromosome")
    assert rc == 0 and mode == 0
    world = _world(backend, 2 if backend == "hub-1gpu" else 8)

    def body(rank, mg, ctx):
        return mg.compress_sharded(ref, tgt, b">sharded chromosome")
    res = T.run_ranks(world, body)
    got, gmode, sharded = res[0]
    assert sharded and gmode == 0 and got == exp
