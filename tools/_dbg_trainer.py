import os, sys, random, torch, warnings, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantization_b200 import QuantizerTrainer, Quantizer, synth
DEV = torch.device("cuda:0")
dim, B = 64, 2048
xs = [synth.synth_x(B, dim, 900 + i).to(DEV) for i in range(4)]
def fresh_indexes(q, x, n):
    f = Quantizer(q.dim, q.codebook_size, q.num_codebooks).to(DEV)
    f.load_state_dict(q.state_dict())
    with torch.no_grad():
        return f._compute_indexes(x, n)
def run():
    def make(use_graph, name):
        torch.manual_seed(3); random.seed(3)
        tr = QuantizerTrainer(dim=dim, bytes_per_frame=2, device=DEV, phase_one_iters=10000, phase_two_iters=10000)
        tr._use_graph = use_graph; tr.two_iter_prob = 0.0
        orig_eager = tr._eager_update
        def eager_update(x, n):
            if tr.cur_iter == 200:
                q = tr.quantizer
                with torch.no_grad():
                    a = q._compute_indexes(x, n).clone()
                    key_before = q._prep_key
                    q._prep_key = None
                    b = q._compute_indexes(x, n).clone()
                    c = fresh_indexes(q, x, n)
                    l0 = [float(v) for v in q.compute_loss(x, n)]
                print(name, "step 200: cached-vs-reprepared differ", int((a != b).any(1).sum()), "reprepared-vs-fresh-module", int((b != c).any(1).sum()),
                      "versions", [p._version for p in q.parameters()], "losses", ["%.5f" % v for v in l0], flush=True)
            return orig_eager(x, n)
        tr._eager_update = eager_update
        return tr
    eager, graphed = make(False, "eager  "), make(True, "graphed")
    for t in (eager, graphed): t.cur_iter = 190
    for i in range(11):
        for t in (eager, graphed): t.step(xs[i % 4])
    torch.cuda.synchronize()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    run()
