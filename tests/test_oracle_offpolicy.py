"""Pin oracle/offpolicy.py against the reference's own TD3.train / DDPG.train outputs (CPU)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import offpolicy as OP
from oracle import onpolicy as O

O_DIM, A_DIM, H = 11, 3, 64
PS, QS = [O_DIM, H, H, A_DIM], [O_DIM + A_DIM, H, H, 1]


def setup(g, twin):
    nets = {"policy": O.unflatten_layers(g["policy_flat0"], PS), "q1": O.unflatten_layers(g["q1_flat0"], QS)}
    nets["target_policy"] = O.unflatten_layers(g["policy_flat0"], PS)
    nets["target_q1"] = O.unflatten_layers(g["q1_flat0"], QS)
    if twin:
        nets["q2"] = O.unflatten_layers(g["q2_flat0"], QS)
        nets["target_q2"] = O.unflatten_layers(g["q2_flat0"], QS)
    adams = {k: O.AdamState(g[k + "_flat0"].size, 1e-3) for k in (["policy", "q1", "q2"] if twin else ["policy", "q1"])}
    mbs = [{k: g["mb_" + k][s] for k in ("observations", "actions", "rewards", "next_observations", "dones")}
           for s in range(int(g["S"]))]
    return nets, adams, mbs


@pytest.mark.parametrize("case,twin", [("td3_small", True), ("ddpg_small", False)])
def test_offpolicy_oracle_matches_reference(case, twin):
    g = load_golden(case)
    nets, adams, mbs = setup(g, twin)
    logs = OP.offpolicy_train(nets, adams, mbs, g["noise"] if twin else None, policy_delay=2 if twin else 1, twin=twin)
    names = ["policy", "q1"] + (["q2"] if twin else [])
    for n in names:
        assert rel_err(O.flatten_layers(nets[n]), g[n + "_flat_final"]) < 1e-5, n
        assert rel_err(O.flatten_layers(nets["target_" + n]), g["target_" + n + "_flat_final"]) < 1e-5, n
    pre = "q-function_1" if twin else "q-function"
    assert abs(np.mean(logs["q1_losses"]) - g["metric:" + pre + "/average_loss"]) < 1e-5
    assert abs(np.mean(np.concatenate(logs["q1_values"])) - g["metric:" + pre + "/avarage_q-value"]) < 1e-6
    assert abs(np.mean(logs["policy_losses"]) - g["metric:policy/average_loss"]) < 1e-6
    assert len(logs["policy_losses"]) == (4 if twin else 7)
