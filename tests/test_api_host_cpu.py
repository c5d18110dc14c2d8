"""Host-side helpers of the Python mirror that need no GPU: the selector's random pattern and LiDAR density (must equal the oracle's / the reference's), the CSR packing of the
batched activation call."""
import numpy as np
import orc


def _api():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api
    return api


def test_random_pattern_and_density_match_the_oracle_helpers():
    api = _api()
    assert np.array_equal(api.random_pattern(64, 48), orc.libc_random_pattern(64, 48))
    for lrud, wh, d in (([4, 634, 74, 188], (640, 192), 600.0), ([10000, -1, 10000, -1], (1200, 360), 500.0), ([12, 1187, 140, 356], (1200, 360), 1500.0)):
        assert api.lidar_density(lrud, wh, d) == orc.lidar_density(lrud, wh, d)


def test_pack_activation_offsets():
    api = _api(); rng = np.random.default_rng(0)
    def seq(nh, npts, nch, ncs):
        pb = np.concatenate([[0], np.cumsum(npts)]); cb = np.concatenate([[0], np.cumsum(ncs)])
        return dict(pt_begin=pb, KRKi=rng.normal(size=(nh, 9)), Kt=rng.normal(size=(nh, 3)), uvid=rng.normal(size=(pb[-1], 3)), cand_begin=cb, cKRKi=rng.normal(size=(nch, 9)), cKt=rng.normal(size=(nch, 3)),
                    cand4=rng.normal(size=(cb[-1], 4)), minActDist=float(nh))
    S = [seq(2, [3, 2], 3, [2, 0, 4]), seq(1, [4], 2, [1, 1]), dict(pt_begin=[0, 2], KRKi=np.zeros((1, 9)), Kt=np.zeros((1, 3)), uvid=np.zeros((2, 3)))]   # the last one: a map only, no candidates
    P = api.packActivation(S)
    assert list(P["hb"]) == [0, 2, 3, 4] and list(P["pb"]) == [0, 3, 5, 9, 11] and list(P["gb"]) == [0, 3, 5, 5] and list(P["cb"]) == [0, 2, 2, 6, 7, 8]
    assert P["A"].shape == (4, 9) and P["U"].shape == (11, 3) and P["cA"].shape == (5, 9) and P["c4"].shape == (8, 4) and list(P["md"]) == [2.0, 1.0, 0.0] and len(P["dec"]) == 8
    assert np.array_equal(P["U"][5:9], S[1]["uvid"].astype(np.float32)) and np.array_equal(P["c4"][6:8], S[1]["cand4"].astype(np.float32))
