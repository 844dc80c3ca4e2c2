import sys, time, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import scipy.io, torch
from amcpy_b200 import synth, ops, matio
from amcpy_b200 import feature_extraction as fe
from amcpy_b200.config import Config, Paths, SignalConfig
with tempfile.TemporaryDirectory() as td:
    cfg = Config(paths=Paths(root=Path(td)), signals=SignalConfig(num_frames=500))
    cfg.paths.ensure_dirs()
    snrs = [float(v) for v in cfg.signals.snr_values.values()]
    x = synth.dataset_device(6, snrs, 500, 2048, torch.device("cuda"), seed=1).view(6, 16, 500, 2048)
    synth.write_all_modulations_mat(cfg.paths.mat_data / cfg.paths.mat_filename, x.cpu().numpy(), cfg.signals.mat_info)
    del x
    fe.run_extraction(cfg)
    T = {}
    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter(); r = fn(*a, **k); T[name] = T.get(name, 0) + time.perf_counter() - t0; return r
        return w
    matio.read_planar = timed("read_planar", matio.read_planar)
    ops.extract_features_host_planar = timed("extract_planar", ops.extract_features_host_planar)
    fe.extract_modulation_planar = timed("extract_modulation_planar(total)", fe.extract_modulation_planar)
    scipy.io.savemat = timed("savemat", scipy.io.savemat)
    t0 = time.perf_counter(); fe.run_extraction(cfg); tot = time.perf_counter() - t0
    print("total", round(tot, 4), {k: round(v, 4) for k, v in T.items()})
    # how much of it is first-touch page faulting of the fresh memory map?  (same mapping, second pass)
    src = fe._MatSource(cfg.paths.mat_data / cfg.paths.mat_filename)
    for rep in range(2):
        t0 = time.perf_counter()
        for m in cfg.signals.modulations_with_noise:
            src.features(cfg.signals.mat_info[m], m, 16, 500, 2048, 0)
        print("pass over one mapping", rep, round(time.perf_counter() - t0, 4))
