"""K3 (avd_score_clips): the per-clip scoring tail on the device against the oracle's NumPy restatement of
video.py:54-83 / fusion.py:16-109 and against the reference's own fuse() outputs in tests/golden/golden.json.
float64 with the reference's operation order: everything is compared with ==."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "these tests need a GPU"
    return t


@pytest.fixture(scope="module")
def eng(torch):
    from avd_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_fuse_goldens_on_device(golden, eng):
    """The 40 fuse() cases produced by the reference's fusion.py, scored by the kernel from the given video dict."""
    from avd_b200 import scoring
    jl = scoring.JobList()
    for c in golden["fuse"]:
        jl.add(video=c["video"], audio=c["audio"], hints=c["hints"])
    l0 = eng.launches
    jobs, sc, vt, fu, bi, pk = scoring.score_jobs(eng, jl)
    assert eng.launches == l0 + 1                                    # one kernel for the whole batch
    for i, c in enumerate(golden["fuse"]):
        assert scoring.fuse_dict(jobs[i], sc[i], fu, bi, pk, c["hints"]) == c["out"], i
        o, L = int(jobs[i]["out_off"]), int(sc[i]["L"])
        if c["video"]["timeline"]:
            assert vt[o:o + L].tolist() == c["video_timeline_after"], i   # fusion.py:21 pads the caller's list
        assert int(sc[i]["n_peaks"]) == len(c["out"]["peaks"])


def _random_clip_records(rng, n, trial):
    from avd_b200._lib import REC_DTYPE
    r = np.zeros(n, REC_DTYPE)
    r["lap_sum"] = rng.integers(-10**6, 10**6, n)
    r["lap_sumsq"] = rng.integers(10**8, 10**12, n)
    if trial % 5 == 0:                                               # flat footage: texture -> 0, ai_susp -> 1
        r["lap_sumsq"] = rng.integers(0, 10**7, n)
        r["lap_sum"] = 0
    r["flow_mean"] = (rng.uniform(0, 12, n) if trial % 3 else rng.uniform(0, 0.3, n)).astype(np.float32)
    r["flow_var"] = rng.uniform(0, 1.2, n).astype(np.float32)
    r["ham_prev"] = rng.choice([0, 0, 5, 100], n)
    r["flags"] = 3
    return r


def test_records_tail_and_fuse_vs_oracle(torch, eng, oracle):
    """Random record batches (0 .. 300 samples per clip: every branch of the pairwise sums, padding and truncation
    of the timeline, with and without audio) scored on the device == the oracle, field by field.  The first record
    of a clip is a foreign one on purpose (flags set): the kernel must ignore its predecessor like video.py:20,26."""
    from avd_b200 import hints as H, scoring
    from avd_b200._lib import REC_DTYPE
    rng = np.random.default_rng(3)
    for trial in range(60):
        recs, specs, jl = [], [], scoring.JobList()
        for ci in range(int(rng.integers(1, 7))):
            n = int(rng.choice([0, 1, 2, 3, 7, 8, 9, 20, 130, 300]))
            r = _random_clip_records(rng, n, trial)
            dur = float(rng.choice([0.0, 0.4, 1.5, 2.5, 5.0, 10.0, 60.0, 200.0]))
            audio = None
            if ci % 3 != 2:
                audio = {"timeline": rng.uniform(0, 1, int(rng.integers(0, 15))).tolist(),
                         "flags_audio": {"speech_ratio": float(rng.choice([0.0, 0.3, 0.9])), "tts_like": float(rng.choice([0.0, 0.7, 0.97]))}}
            hn = H.compute_hints({"width": 1280, "height": 720, "fps": 30.0, "bit_rate": int(rng.choice([0, 900000, 3500000, 9000000]))})
            jl.add(rec_first=len(recs), rec_count=n, npix=1280 * 720, tlen=scoring.timeline_length(dur),
                   audio=copy.deepcopy(audio), hints=hn)
            specs.append((len(recs), n, dur, audio, hn))
            recs.extend(r.tolist())
        R = np.array(recs, dtype=REC_DTYPE) if recs else np.zeros(1, REC_DTYPE)
        d_rec = torch.from_numpy(R.view(np.uint8).reshape(-1)).cuda()
        jobs, sc, vt, fu, bi, pk = scoring.score_jobs(eng, jl, d_rec.data_ptr(), len(recs))
        for ci, (f0, n, dur, audio, hn) in enumerate(specs):
            r = R[f0:f0 + n]
            orecs = []
            for i in range(n):
                d = {"lap_sum": int(r["lap_sum"][i]), "lap_sumsq": int(r["lap_sumsq"][i]), "ham": int(r["ham_prev"][i]) if i else -1}
                if i > 0:
                    d["flow_mean"], d["flow_var"] = float(r["flow_mean"][i]), float(r["flow_var"][i])
                orecs.append(d)
            want_v = oracle.video_result(orecs, 720, 1280, 30.0, dur)
            want_f = oracle.fuse(copy.deepcopy(audio), want_v, hn, raw=True) if audio is not None else None
            got_v = scoring.video_dict(jobs[ci], sc[ci], vt, 1280, 720, 30.0)
            assert got_v == want_v, (trial, ci, n, dur)
            if audio is not None:
                assert scoring.fuse_dict(jobs[ci], sc[ci], fu, bi, pk, hn, with_raw=True) == want_f, (trial, ci, n, dur)


def test_clip_batch_frames_to_verdicts(synth, eng):
    """Frames in, verdicts out (scoring.score_clip_batch: K1 + K2 + K3, records never leave the device) == the host
    tail on the records of the same frames (video.analyze_frames_batch + fusion.fuse)."""
    from avd_b200 import fusion, hints as H, scoring, video
    clips, metas, audios = [], [], []
    for ci in range(5):
        c = synth.Clip(60, 360, 640, seed=70 + ci, fps=30.0)
        n = [4, 1, 7, 3, 6][ci]
        clips.append([c[i * 15 % 60] for i in range(n)])
        metas.append({"fps": 30.0, "width": 640, "height": 360, "duration": [2.0, 3.0, 2.5, 9.0, 0.0][ci], "bit_rate": 300000 * ci})
        audios.append(fusion.neutral_audio(scoring.timeline_length(metas[-1]["duration"])) if ci != 2 else
                      {"timeline": [0.9, 0.1, 0.4, 0.8, 0.7], "flags_audio": {"speech_ratio": 0.6, "tts_like": 0.97}})
    hints = [H.compute_hints(m) for m in metas]
    want_v = video.analyze_frames_batch(clips, metas, engine=eng)
    a_host = copy.deepcopy(audios)
    want_f = [fusion.fuse(a_host[i], want_v[i], hints[i], with_raw=True) for i in range(5)]
    got = scoring.score_clip_batch(eng, clips, metas, audios=audios, hints=hints, with_raw=True)
    for i in range(5):
        assert got[i]["video"] == want_v[i], i                       # want_v[i]["timeline"] was padded in place by fuse
        assert got[i]["fusion"] == want_f[i], i
        assert audios[i]["timeline"] == a_host[i]["timeline"], i
    only_video = scoring.score_clip_batch(eng, clips, metas)
    ref_v = video.analyze_frames_batch(clips, metas, engine=eng)
    assert [o["video"] for o in only_video] == ref_v and all(o["fusion"] is None for o in only_video)


def test_score_clips_argument_errors(eng):
    from avd_b200 import scoring
    from avd_b200._lib import AvdError
    jl = scoring.JobList()
    jl.add(rec_first=0, rec_count=4, npix=100, tlen=2)
    with pytest.raises(AvdError):
        scoring.score_jobs(eng, jl, None, 4 + 10**6)                 # more records than the last host run left
    jl = scoring.JobList()
    jl.add(rec_first=0, rec_count=0, npix=0, tlen=3)                  # a clip with no samples: [0.5]*tlen, zero summary
    jobs, sc, vt, fu, bi, pk = scoring.score_jobs(eng, jl, None, 0)
    assert vt[:3].tolist() == [0.5, 0.5, 0.5] and float(sc[0]["texture_var"]) == 0.0 and int(sc[0]["vlen"]) == 3


def test_config3_scale_batch(torch, eng, oracle):
    """BASELINE configs[3]: 256 clips x 10 samples in ONE launch, every clip against the oracle's tail + fuse."""
    from avd_b200 import fusion, hints as H, scoring
    from avd_b200._lib import REC_DTYPE
    rng = np.random.default_rng(11)
    R = np.concatenate([_random_clip_records(rng, 10, ci) for ci in range(256)])
    hn = H.compute_hints({"width": 1280, "height": 720, "fps": 30.0, "bit_rate": 2_000_000})
    jl = scoring.JobList()
    for ci in range(256):
        jl.add(rec_first=10 * ci, rec_count=10, npix=1280 * 720, tlen=5, audio=fusion.neutral_audio(5), hints=hn)
    d_rec = torch.from_numpy(R.view(np.uint8).reshape(-1)).cuda()
    l0 = eng.launches
    jobs, sc, vt, fu, bi, pk = scoring.score_jobs(eng, jl, d_rec.data_ptr(), len(R))
    assert eng.launches == l0 + 1
    for ci in range(256):
        r = R[10 * ci:10 * ci + 10]
        orecs = [{"lap_sum": int(r["lap_sum"][i]), "lap_sumsq": int(r["lap_sumsq"][i]), "ham": int(r["ham_prev"][i]) if i else -1,
                  **({"flow_mean": float(r["flow_mean"][i]), "flow_var": float(r["flow_var"][i])} if i else {})} for i in range(10)]
        want_v = oracle.video_result(orecs, 720, 1280, 30.0, 5.0)
        want_f = oracle.fuse(fusion.neutral_audio(5), want_v, hn, raw=True)
        assert scoring.video_dict(jobs[ci], sc[ci], vt, 1280, 720, 30.0) == want_v, ci
        assert scoring.fuse_dict(jobs[ci], sc[ci], fu, bi, pk, hn, with_raw=True) == want_f, ci
