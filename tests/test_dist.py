"""not-gpu: the N>1 path (image sharding + final metric all-reduce) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from shallow_ntc_b200 import parallel, synthetic, build_config


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _fake_image_metrics(i):
  return np.array([30.0 + 0.1 * i, 50.0 - i, 1000.0 * i, 10.0 * i, 1.0])


def _worker(rank, world, port, n_images, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  lo, hi = parallel.shard_range(n_images, rank, world)
  # this rank's latents are exactly the slice of the global seeded list
  m = build_config("two_layer_syn")
  zs, ys = m.latent_shapes(hi - lo, 64, 64)
  z, qy = synthetic.make_latents(zs, ys, first_index=lo)
  sums = sum((_fake_image_metrics(i) for i in range(lo, hi)), np.zeros(5))
  total = parallel.reduce_metric_sums(dist, sums)
  tmax = parallel.max_over_ranks(dist, [1.0 + rank, 5.0 - rank])
  q.put((rank, lo, hi, total.tolist(), tmax.tolist(), float(qy.sum()), float(z.sum())))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_sharding_and_metric_reduce():
  world, n_images = 2, 5
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in range(world))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]                  # contiguous blocks, every image once
  serial = sum((_fake_image_metrics(i) for i in range(n_images)), np.zeros(5))
  for r in res:
    assert np.allclose(r[3], serial)                                      # both ranks hold the global sums
    assert r[4] == [2.0, 5.0]
  assert parallel.mean_metrics(serial)["n_images"] == n_images
  # sharding never changes the data: the union of the shards is the single-process batch
  m = build_config("two_layer_syn")
  zs, ys = m.latent_shapes(n_images, 64, 64)
  z, qy = synthetic.make_latents(zs, ys)
  assert abs(sum(r[5] for r in res) - float(qy.sum())) < 1e-3 and abs(sum(r[6] for r in res) - float(z.sum())) < 1e-3


def test_single_process_passthrough():
  s = parallel.reduce_metric_sums(None, [1, 2, 3, 4, 5])
  assert s.tolist() == [1, 2, 3, 4, 5] and parallel.max_over_ranks(None, [3.0]).tolist() == [3.0]
