"""
CPU, world_size = 2 over gloo: the multi-GPU protocol of bench.py / SURVEY.md section 8e -- point-range sharding of one
MSM, all-gather of the 224-byte partial blocks, fold on one rank -- with the kernel-logic emulator standing in for
the two GPUs.  Also checks the in-library multi-device fold entry point.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path[:0] = [%(root)r, %(root)r + '/oracle', %(root)r + '/tests']
import numpy as np, torch, torch.distributed as dist
import oracle_c as oc, parity_cases as pc
from halo2_scaffold_b200._lib import Lib
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
L = Lib(%(emu)r, allow_emulator=True); L.init(1)
n = 3000
s, P = oc.random_fr(41, n), oc.gen_points(42, n)          # every rank derives the same global problem ...
lo, hi = n * rank // world, n * (rank + 1) // world       # ... and owns one point range
d_s, d_p, d_b = L.dev_alloc(0, (hi - lo) * 32), L.dev_alloc(0, (hi - lo) * 64), L.dev_alloc(0, 224)
L.h2d(0, d_s, s[lo:hi]); L.h2d(0, d_p, P[lo:hi])
L.msm_dev_partial(0, d_s, d_p, hi - lo, d_b); L.dev_sync(0)
block = np.zeros(28, dtype=np.uint64); L.d2h(0, block, d_b)
mine = torch.from_numpy(block.view(np.int64).copy())
gathered = [torch.zeros(28, dtype=torch.int64) for _ in range(world)]
dist.all_gather(gathered, mine)
if rank == 0:
    blocks = np.stack([g.numpy().view(np.uint64) for g in gathered])
    got = pc.affine_of(oc, L.msm_fold_partials(blocks))
    want = pc.affine_of(oc, oc.best_multiexp(s, P))
    assert (got == want).all()
    print('FOLD_OK')
dist.barrier(); dist.destroy_process_group()
"""


def test_point_range_sharding_two_ranks(emu, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "emu": emu.path})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", H2B_EMU_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FOLD_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])


WORKER_ROWS = r"""
import os, sys
sys.path[:0] = [%(root)r, %(root)r + '/oracle', %(root)r + '/tests']
import numpy as np, torch, torch.distributed as dist
import bn254 as o, oracle_c as oc, parity_cases as pc
from halo2_scaffold_b200._lib import Lib
from halo2_scaffold_b200 import evaluation as ev
from halo2_scaffold_b200.domain import fr_to_words
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
L = Lib(%(emu)r, allow_emulator=True); L.init(1)
ek, k, seed = 7, 5, 12
size, rot_scale = 1 << ek, 1 << (ek - k)
polys, lookup_exprs, nf, na = pc.standard_plonk_like(1)
E = ev.Evaluator(polys, lookup_exprs)
col = lambda tag, j=0: oc.random_fr(seed * 1000 + tag + j, size)              # every rank derives the same global columns ...
perm_cols = [("advice", j) for j in range(na)] + [("fixed", 0)]
kw = dict(size=size, rot_scale=rot_scale, fixed=[col(0, j) for j in range(nf)], advice=[col(100, j) for j in range(na)], instance=[col(200)],
          challenges=np.zeros((0, 4), dtype=np.uint64), y=col(300)[0], beta=col(300)[1], gamma=col(300)[2], theta=col(300)[3], l0=col(400), l_last=col(401),
          l_active_row=col(402),
          permutation=dict(product_cosets=[col(500, j) for j in range(2)], columns=perm_cols, cosets=[col(600, j) for j in range(4)], chunk_len=2,
                           last_rotation=-6, delta=col(300)[4], zeta=col(300)[5], extended_omega=fr_to_words(o.omega_for(ek))),
          lookups=[dict(product_coset=col(700), permuted_input_coset=col(701), permuted_table_coset=col(702))], lib=L)
rows = size // world                                                          # ... and evaluates one row range from its slices (+ halo)
mine = E.evaluate_h(shard=(rank * rows, rows, 6 * rot_scale), **kw)
gathered = [torch.zeros(rows * 4, dtype=torch.int64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(mine.view(np.int64).reshape(-1).copy()))
if rank == 0:
    got = np.concatenate([g.numpy().view(np.uint64).reshape(-1, 4) for g in gathered])
    # against the ORACLE's three sequential loops over the whole domain (not against an unsharded run of the same library)
    tup = lambda g: (lambda a: (a.constants, a.rotations, a.calculations, a.parts, a.n_intermediates))(g.arrays())
    ch, pm, lk = np.zeros((0, 4), dtype=np.uint64), kw['permutation'], kw['lookups'][0]
    by_type = {'advice': kw['advice'], 'fixed': kw['fixed'], 'instance': kw['instance']}
    want = oc.evaluate_graph(tup(E.custom_gates), kw['fixed'], kw['advice'], kw['instance'], ch, kw['beta'], kw['gamma'], kw['theta'], kw['y'],
                             np.zeros((size, 4), dtype=np.uint64), rot_scale)
    want = oc.evaluate_h_permutation(want, rot_scale, pm['product_cosets'], [by_type[t][i] for t, i in perm_cols], pm['cosets'], 2, -6, kw['l0'], kw['l_last'],
                                     kw['l_active_row'], kw['beta'], kw['gamma'], kw['y'], pm['delta'], pm['zeta'], pm['extended_omega'])
    want = oc.evaluate_h_lookup(tup(E.lookups[0]), kw['fixed'], kw['advice'], kw['instance'], ch, kw['beta'], kw['gamma'], kw['theta'], kw['y'], want, rot_scale,
                                lk['product_coset'], lk['permuted_input_coset'], lk['permuted_table_coset'], kw['l0'], kw['l_last'], kw['l_active_row'])
    assert (got == want).all()
    print('ROWS_OK')
dist.barrier(); dist.destroy_process_group()
"""


def test_row_sharded_evaluate_h_two_ranks(emu, tmp_path):
    # evaluate_h sharded by rows: two ranks, each with its slice + halo of every column; the gathered rows equal
    # the oracle's evaluation of the whole domain
    script = tmp_path / "worker_rows.py"
    script.write_text(WORKER_ROWS % {"root": ROOT, "emu": emu.path})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537", H2B_EMU_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29537", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ROWS_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--k", "12", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["value"] > 0
