#! /usr/bin/python3
# Same flow as the reference's examples/python/ggnn_pytorch_gpu_data.py and ggnn_pytorch.py, written against `import ggnn`
# (run with PYTHONPATH=<repo>/compat/python:<repo>): base and query tensors that already live on the GPU are used in
# place, results stay on the GPU; then ground truth + evaluation.
import ggnn
import torch

ggnn.set_log_level(1)

base = torch.rand((10_000, 128), dtype=torch.float32, device='cuda')
query = torch.rand((10_000, 128), dtype=torch.float32, device='cuda')

my_ggnn = ggnn.GGNN()
my_ggnn.set_base(base)
my_ggnn.set_return_results_on_gpu(True)

measure = ggnn.DistanceMeasure.Euclidean
my_ggnn.build(k_build=24, tau_build=0.5, measure=measure)

k_query: int = 10
indices, dists = my_ggnn.query(query, k_query, 0.64, 400, measure)
assert indices.is_cuda and dists.is_cuda

gt_indices, gt_dists = my_ggnn.bf_query(query, k_gt=k_query, measure=measure)
evaluator = ggnn.Evaluator(base, query, gt_indices, k_query=k_query)
print(evaluator.evaluate_results(indices))
print('indices:', indices[:2], '\n squared dists:', dists[:2], '\n')
