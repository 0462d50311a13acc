import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, synth, stratum_dsp_b200 as S
SR, N, nt = 44100, 7_938_000, 64
params = np.array([[c.bpm, c.tonic, c.minor, c.phase_frac, c.chord_amp] for c in (synth.c2_params(i) for i in range(nt))], np.float32)
buf = torch.empty(nt * N, dtype=torch.float32, device="cuda")
S.synth_batch(buf.data_ptr(), nt, N, SR, params)
off = np.arange(nt + 1, dtype=np.uint64) * N
for name, cfg in (("default", None), ("hpss_onsets", S.AnalysisConfig(enable_hpss_onsets=1)), ("perc_fallback", S.AnalysisConfig(enable_tempogram_percussive_fallback=1))):
    S.stage_timing(True); S.stage_times(reset=True)
    for _ in range(2):
        r = S.analyze_batch_device(buf.data_ptr(), off, [SR] * nt, cfg, convert=False); ms = S.last_call_device_ms(); S.free_results(r)
    st = S.stage_times(reset=True)
    print(name, "ms/track", round(ms / nt, 3), "hpss stage ms/track", round(st.get("hpss", 0) / 2 / nt, 3), round(st.get("percussive_fallback", 0) / 2 / nt, 3))
