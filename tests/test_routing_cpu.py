"""CPU check of the decision-injection test infrastructure (tests/routing.py): with the oracle's OWN decisions injected
the routed oracle is the oracle (outputs and second-order quantities), and the decision report shows no flips."""
import torch

from oracle import port
from tests import routing


def test_routed_oracle_with_own_decisions_is_the_oracle():
    torch.manual_seed(3)
    omodel = port.make_cifar10net(2, 4, use_tanh=True)
    g = torch.Generator().manual_seed(4)
    X = torch.randn(5, 2, 28, 28, generator=g)
    dec = routing.oracle_decisions(omodel, X)
    assert [k for k, _ in dec] == ['relu', 'pool', 'relu', 'pool', 'relu', 'pool']
    routed = routing.routed_oracle(omodel, dec)
    with torch.no_grad():
        assert torch.equal(routed(X), omodel(X))
    lh = port.CategoricalLh()
    assert torch.allclose(port.diag_ggn_batch(routed, lh, X), port.diag_ggn_batch(omodel, lh, X), rtol=1e-6, atol=1e-9)
    for qa, qb in zip(port.kflr_batch(routed, lh, X), port.kflr_batch(omodel, lh, X)):
        for a, b in zip(qa, qb):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)
    rep = routing.decision_report(omodel, dec, X)
    assert rep['relu_flips'] == 0 and rep['pool_flips'] == 0 and rep['relu_total'] > 0 and rep['pool_total'] > 0
    # a deliberately wrong decision is reported with its margin
    kind, mask = dec[0]
    mask = mask.clone()
    mask[0, 0, 0, 0] = ~mask[0, 0, 0, 0]
    rep = routing.decision_report(omodel, [(kind, mask)] + dec[1:], X)
    assert rep['relu_flips'] == 1 and rep['relu_worst_margin'] > 0
