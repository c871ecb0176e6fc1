import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from oracle import port
from preds_b200 import Laplace, CategoricalLh, MLPS
from preds_b200.predictives import glm_moments

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())

torch.manual_seed(71)
model = MLPS(6, [16, 12], 3, 'tanh')
g = torch.Generator().manual_seed(15)
X = torch.randn(96, 6, generator=g); y = torch.randint(3, (96,), generator=g)
Xte = torch.randn(10, 6, generator=g); eps = torch.randn(7, 10, 3, generator=g)
omodel = port.make_mlps(6, [16, 12], 3, 'tanh'); port.set_theta(omodel, port.get_theta(model))
model = model.cuda()
batches = [(X[:64], y[:64]), (X[64:], y[64:])]
post = port.infer(omodel, 0.8, port.CategoricalLh(), batches, 'full')
for eng in (1, 2):
    lap = Laplace(model, 0.8, CategoricalLh()); lap.ggn_engine = eng
    try:
        lap.infer(batches, cov_type='full')
    except Exception as e:
        print('engine', eng, 'FAILED', repr(e)); continue
    print('engine', eng, 'H err', rel(lap.precision.cpu(), post['precision']))
    d = (lap.precision.cpu() - post['precision']).abs()
    print('   worst idx', divmod(int(d.argmax()), d.shape[1]), 'sym err', float((lap.precision - lap.precision.T).abs().max()))
    Sig_o = post['Sigma_chol'] @ post['Sigma_chol'].T
    print('   Sigma err', rel(lap.Sigma.cpu(), Sig_o), 'cond', float(torch.linalg.cond(post['precision'].double())))
    fmu, fvar = glm_moments(Xte, model, lap.mu, lap.Sigma.contiguous())
    ofmu, ofvar = port.glm_moments(omodel, Xte, post)
    print('   fmu err', rel(fmu.cpu(), ofmu), 'fvar err', rel(fvar.cpu(), ofvar))
    # fvar with the ORACLE Sigma on device: isolates the f_var kernel from the factorisation
    fmu2, fvar2 = glm_moments(Xte, model, lap.mu, Sig_o.cuda().contiguous())
    print('   fvar err with oracle Sigma', rel(fvar2.cpu(), ofvar))
    out = lap.predictive_samples_glm(Xte, n_samples=7, eps=eps).cpu()
    ref = port.glm_predictive(omodel, port.CategoricalLh(), Xte, post, eps)
    print('   glm err', rel(out, ref))
