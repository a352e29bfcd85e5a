"""configs[0]: the reference's own driver loop (robot-learning.py:66-101) run with the drop-in `Environment` / `Robot`:
same numpy seed, same torch seed, 130 ticks = three demonstrations, the leaving-demo reset, two episodes and two TD3
updates.  Action types, done flags, replay indices and the whole numpy RNG consumption are bit-exact; states, actions and
rewards agree to float32 round-off of the float64 reference; losses of the updates to 1e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_driver_loop_trace_vs_reference(pkg, env_golden, trace_golden):
    g, e = trace_golden, env_golden
    c = pkg.constants
    np.random.seed(1707366464)                             # robot-learning.py:19
    torch.manual_seed(0)
    environment = pkg.Environment(maps=(e["speed"], e["angle"]))
    state = environment.reset()
    robot = pkg.Robot(environment.goal_state)
    assert (environment.goal_state == g["goal"]).all() and (environment.robot_init_region == g["region"]).all()
    assert (state == g["state0"]).all()
    # the networks consume torch's RNG exactly like the reference's constructors: identical initial weights
    for net, k in ((0, "actor"), (1, "critic1"), (2, "critic2")):
        assert (robot.td3_agent.flat(net).cpu().numpy() == g["w_" + k]).all(), k
    # target-policy noise comes from torch's (unseeded, in the reference) generator: inject the recorded draws
    real_update = robot.td3_agent.td3_update
    upd = {"n": 0, "losses": []}

    def update_with_recorded_noise(memory):
        z = torch.from_numpy(g["update_noise"][upd["n"] * 100:(upd["n"] + 1) * 100]).cuda()
        upd["losses"].append(real_update(memory, noise=z))
        upd["n"] += 1
    robot.td3_agent.td3_update = update_with_recorded_noise

    demos = steps = 0
    demos_bought = resets_bought = steps_bought = 0
    for tick in range(g["types"].shape[0]):
        money = c.STARTING_MONEY - (demos_bought * c.COST_PER_DEMO + resets_bought * c.COST_PER_RESET + steps_bought * c.COST_PER_STEP)
        action_type = robot.get_next_action_type(state, money)
        assert {"step": 0, "demo": 1, "reset": 2}[action_type] == g["types"][tick], tick
        if action_type == "reset":
            state = environment.reset()
            resets_bought += 1
            assert (state == g["states"][tick]).all()      # reset draws are bit-exact
        elif action_type == "demo":
            ds, da = environment.get_demonstration()       # consumes numpy's stream like the reference's planner ...
            assert ds.shape == (200, 2) and np.abs(ds[0] - g["demo_states"][demos][0]).max() < 1e-5
            # ... its float32 paths differ from the float64 reference in the last digits; continue from the reference's demo
            robot.process_demonstration(g["demo_states"][demos], g["demo_actions"][demos], money)
            demos += 1
            demos_bought += 1
        else:
            action = robot.get_next_action_training(state, money)
            next_state = environment.step(action)
            robot.process_transition(state, action, next_state, money)
            tol = 2e-4 if upd["n"] == 0 else 5e-3         # after an update the actors agree to the 1e-3 learner tolerance
            np.testing.assert_allclose(action, g["actions"][tick], rtol=0, atol=tol, err_msg="tick %d" % tick)
            np.testing.assert_allclose(next_state, g["states"][tick], rtol=0, atol=tol, err_msg="tick %d" % tick)
            np.testing.assert_allclose(float(robot._reward64[0]), g["step_rewards"][steps], rtol=0, atol=20 * tol)
            assert bool(robot._done[0]) == bool(g["step_dones"][steps])
            state = g["states"][tick].copy()               # teacher-forced: continue from the reference's state
            environment.robot_state = state
            steps += 1
            steps_bought += 1
    assert upd["n"] == int(g["n_updates"]) and len(robot.memory) == int(g["replay_len"])
    closs = torch.cat([l[0] for l in upd["losses"]]).cpu().numpy()
    aloss = torch.cat([l[1] for l in upd["losses"]]).cpu().numpy()
    np.testing.assert_allclose(closs, g["critic_losses"], rtol=2e-3)
    np.testing.assert_allclose(aloss, g["actor_losses"], rtol=2e-3)
    # every numpy draw of the run (start/goal, CEM planner, augmentation, exploration noise, 300 replay index sets) was consumed
    # exactly as by the reference
    assert np.random.uniform() == float(g["final_uniform"])
