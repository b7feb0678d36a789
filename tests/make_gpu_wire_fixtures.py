"""Produces tests/golden/gpu_wire_*.{npz,npy,pkl} ON THE GPU BOX: files written by the CUDA path in the reference's wire formats,
committed so that tests/test_wire_formats_reference_readers.py can feed them to the reference's own inspection scripts in the
container that holds /root/reference (the GPU box does not).  Usage: python tests/make_gpu_wire_fixtures.py OUT_DIR"""
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out, only_legacy=False):
    from ffm_b200 import workloads
    from ffm_b200.model.ffm_core import FloorFieldModel as Core
    from ffm_b200.model.ffm_learning_core import FloorFieldModel as Mcq
    from ffm_b200.model.ffm_unified import FloorFieldModelUnified
    os.makedirs(out, exist_ok=True)
    m = workloads.room_map(12, 12)
    with tempfile.TemporaryDirectory() as tmp:
        p32, p64 = os.path.join(tmp, "sff32.npy"), os.path.join(tmp, "sff64.npy")
        np.save(p32, workloads.sff_room(m, "neumann"))
        np.save(p64, workloads.sff_room(m, "neumann").astype(np.float64))
        # main.py:42-52 -> positions.npy
        np.random.seed(1)
        core = Core(m, p32, 25, {"seed": 101, "neighborhood": "neumann"})
        log = []
        while core.positions.shape[0] > 0:
            core.step()
            log.append(np.copy(core.positions))
        np.save(os.path.join(out, "gpu_wire_positions.npy"), np.array(log, dtype=object))
        # run_actor_only_training.py:205-218 -> trajectory_N*.npz
        params = dict(k_S=10, k_D=1, alpha_v=0.01, gamma=0.99, exit_reward=100.0, step_penalty=-1.0, collision_penalty=-1.0,
                      neighborhood="neumann", block_size=1, seed=202)
        uni = FloorFieldModelUnified(m, p32, 12, learning_mode="critic_only", params=params)
        uni.reset(exit_pos=(0, 6), radius=9)
        steps, traj = uni.run(max_steps=200, return_trajectory=True)
        np.savez_compressed(os.path.join(out, "gpu_wire_trajectory_N12_ep00001_total00001.npz"), positions=traj, episode=1, N=12,
                            total_episode=1, steps=steps)
        # run_unified_critic_training.py:290-299 -> V pickle
        with open(os.path.join(out, "gpu_wire_V.pkl"), "wb") as f:
            pickle.dump(uni.get_v_table(), f)
        # ffm_learning_core.py:364-367 -> Q.pkl
        mcq = Mcq(m, p64, 20, {"seed": 303, "max_steps": 80})
        for beta in (1.0, 0.5):
            mcq.reset()
            while mcq.positions.shape[0] > 0:
                mcq.step(beta=beta)
        mcq.save_Q(os.path.join(out, "gpu_wire_Q.pkl"))
    with tempfile.TemporaryDirectory() as tmp:
        # legacy drivers: run_critic_training.py:219-226 -> V_integrated_total*.pkl, run_actor_only_training.py:293-299 -> H_actor_N*.pkl
        from ffm_b200.model.ffm_ac_core import FloorFieldModel as LegacyCritic
        from ffm_b200.model.ffm_actor_only import FloorFieldModelActorOnly
        p32 = os.path.join(tmp, "sff32.npy")
        np.save(p32, workloads.sff_room(m, "neumann"))
        np.random.seed(2)
        critic = LegacyCritic(m, p32, 20, {"seed": 404, "neighborhood": "neumann", "block_size": 5, "step_penalty": -1.0})
        for ep in range(3):
            if ep:
                critic.reset()
            critic.run(max_steps=100)
        vp = os.path.join(out, "gpu_wire_legacy_V.pkl")
        with open(vp, "wb") as f:
            pickle.dump(critic.get_v_table(), f)
        actor = FloorFieldModelActorOnly(m, p32, 1, pretrained_v_path=vp, params={"seed": 505, "neighborhood": "neumann", "step_penalty": -1.0})
        actor.set_epsilon(0.2)
        for ep in range(40):
            if ep:
                actor.reset()
            actor.run(max_steps=100)
        with open(os.path.join(out, "gpu_wire_legacy_H_actor.pkl"), "wb") as f:
            pickle.dump(actor.get_h_table(), f)
    print("written:", sorted(os.listdir(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "wire"))
