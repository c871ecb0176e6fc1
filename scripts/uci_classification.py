"""The UCI classification driver's inference chain (experiments/classification.py:63-176 of the reference, minus the
Bayes-by-backprop baseline) on synthetic satellite-shaped data (SURVEY.md 8 cfg 2: D=36, K=6, N=4504/965,
SiMLP 2x50 tanh, P=4706), through the product package only.  Prints per-stage wall times (CUDA-synchronised).

    python scripts/uci_classification.py [--epochs 200] [--samples 1000] [--refine-epochs 50]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))

from preds_b200 import LaplaceGGN, get_diagonal_ggn, CategoricalLh  # noqa: E402
from preds_b200.models import SiMLP  # noqa: E402
from preds_b200.predictives import linear_sampling_predictive, nn_sampling_predictive  # noqa: E402
from preds_b200.refine import laplace_refine, vi_refine, vi_diag_refine  # noqa: E402
from preds_b200.utils import classification_metrics  # noqa: E402
from torch.nn.utils import parameters_to_vector  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=200)
    ap.add_argument('--samples', type=int, default=1000)
    ap.add_argument('--refine-epochs', type=int, default=50)
    ap.add_argument('--n-train', type=int, default=4504)
    ap.add_argument('--n-test', type=int, default=965)
    a = ap.parse_args()
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(15)
    D, K = 36, 6
    centers = 1.5 * torch.randn(K, D, generator=g)

    def make(n):
        y = torch.randint(K, (n,), generator=g)
        return (centers[y] + torch.randn(n, D, generator=g)).to(dev), y.to(dev)

    Xtr, ytr = make(a.n_train)
    Xte, yte = make(a.n_test)
    torch.manual_seed(71)
    model = SiMLP(D, K, 2, 50, activation='tanh').to(dev)
    lh = CategoricalLh()
    prior_prec = 1.0
    times, res = {}, {}

    def stage(name, fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        times[name] = round(time.perf_counter() - t, 4)
        return out

    opt = LaplaceGGN(model, lr=1e-2, prior_prec=prior_prec)

    def train():
        for _ in range(a.epochs):
            def closure():
                model.zero_grad()
                return lh.log_likelihood(ytr, model(Xtr))
            opt.step(closure)
    stage('train_map', train)
    stage('post_process', lambda: opt.post_process(model, lh, [(Xtr, ytr)]))
    theta_star = parameters_to_vector(model.parameters()).detach()
    Sigmad, Sigma_chold = get_diagonal_ggn(opt)
    Sigma_chol = opt.state['Sigma_chol']

    def evaluate(name, p):
        res[name] = {k: round(v, 5) for k, v in classification_metrics(p, yte, lh, bins=10).items()}

    evaluate('map', lh.inv_link(model(Xte).detach()))
    for name, fn, chol in (('glm', linear_sampling_predictive, Sigma_chol), ('glmd', linear_sampling_predictive, Sigma_chold),
                           ('nn', nn_sampling_predictive, Sigma_chol), ('nnd', nn_sampling_predictive, Sigma_chold)):
        p = stage('pred_' + name, lambda: fn(Xte, model, lh, theta_star, chol, mc_samples=a.samples).mean(0))
        evaluate(name, p)
    m, S_chol, S_chold, _ = stage('laplace_refine', lambda: laplace_refine(model, Xtr, ytr, lh, prior_prec,
                                                                        n_epochs=a.refine_epochs * 4))
    evaluate('glmLap', linear_sampling_predictive(Xte, model, lh, m, S_chol, mc_samples=a.samples).mean(0))
    m, S_chol, losses = stage('vi_refine', lambda: vi_refine(model, opt, Xtr, ytr, lh, n_epochs=a.refine_epochs))
    evaluate('glmVI', linear_sampling_predictive(Xte, model, lh, m, S_chol, mc_samples=a.samples).mean(0))
    res['elbo_glm'] = round(losses[-1], 3)
    m, S_chold, losses = stage('vi_diag_refine', lambda: vi_diag_refine(model, opt, Xtr, ytr, lh, n_epochs=a.refine_epochs))
    evaluate('glmVId', linear_sampling_predictive(Xte, model, lh, m, S_chold, mc_samples=a.samples).mean(0))
    print(json.dumps({'config': vars(a), 'seconds': times, 'metrics': res}))


if __name__ == '__main__':
    main()
