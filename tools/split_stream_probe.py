"""Probe: C3 (8 dialogues) as ONE handle on all SMs vs TWO handles of 4 dialogues on disjoint SM halves on two streams
(tensor-bound GEMMs of one half overlap the MUFU-bound attention of the other; flatter power draw).
python tools/split_stream_probe.py [sms_a]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covomix_b200  # noqa: E402,F401
from covomix_b200 import synthetic as syn  # noqa: E402
from covomix_b200.flow import B200FlowSampler  # noqa: E402


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    cfg = syn.VOMIX
    sd = syn.synthetic_flow_state_dict(cfg, 1234)
    ids, cond, y0, _ = syn.synthetic_flow_inputs(cfg, 8, 1650, prompt=150, seed=30)
    ids, cond, y0 = ids.to(dev), cond.to(dev), y0.to(dev)
    n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    full = B200FlowSampler(sd, cfg, dev, torchdiffeq_ode_method="euler", ode_step_size=1 / 64)
    t_full = timed(lambda: full.sample(phoneme_ids=ids, cond=cond, cond_scale=0.7, y0=y0))
    print(f"one handle, {n_sms} SMs, B=8: {t_full:.1f} ms")
    full.close()
    for sms_a in [int(a) for a in sys.argv[1:]] or [n_sms // 2]:
        a = B200FlowSampler(sd, cfg, dev, torchdiffeq_ode_method="euler", ode_step_size=1 / 64, sm_limit=sms_a)
        b = B200FlowSampler(sd, cfg, dev, torchdiffeq_ode_method="euler", ode_step_size=1 / 64, sm_limit=n_sms - sms_a)
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def split():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                m1 = a.sample(phoneme_ids=ids[:4], cond=cond[:4], cond_scale=0.7, y0=y0[:4])
            with torch.cuda.stream(s2):
                m2 = b.sample(phoneme_ids=ids[4:], cond=cond[4:], cond_scale=0.7, y0=y0[4:])
            cur.wait_stream(s1)
            cur.wait_stream(s2)
            return m1, m2

        t = timed(split)
        print(f"two handles, {sms_a} + {n_sms - sms_a} SMs, 2 x B=4 on two streams: {t:.1f} ms")
        a.close()
        b.close()


if __name__ == "__main__":
    main()
