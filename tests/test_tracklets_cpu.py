"""Host-side half of the tracklet-level entry point: the pair table must describe exactly the pairs, lengths
and frame offsets that the data loader's pair construction (restated in synth.synthetic_video) produces."""
import pytest
import torch

from vrdone_b200 import MaskVRD, synth


@pytest.mark.parametrize("name,kw", [("vidvrd", {}), ("vidor", dict(n_tracklets=12, n_frames=700)), ("vidor_x", dict(n_tracklets=8, n_frames=500))])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pair_table_matches_loader_pairs(name, kw, seed):
    cfg = synth.load_config(name)
    st = cfg["dataset_config"]["feat_stride"]
    trk = synth.synthetic_tracklet_video(cfg, seed, **kw)
    vid = synth.synthetic_video(cfg, seed, **kw)
    keep, L, s_off, o_off = MaskVRD.pair_table(trk["traj_durations"].numpy(), trk["sids"].numpy(), trk["oids"].numpy(), st, 0, 0)
    assert keep.all()
    assert L.tolist() == [int(f.shape[1]) for f in vid["so_features_list"]]
    assert torch.equal(trk["sids"], vid["sids"]) and torch.equal(trk["oids"], vid["oids"])
    nv = cfg["model_config"]["visual_dim"]
    for i in range(0, len(L), max(1, len(L) // 7)):
        s, o = int(trk["sids"][i]), int(trk["oids"][i])
        assert torch.equal(trk["visual_features_list"][s][s_off[i]::st][:L[i]].t(), vid["so_features_list"][i][:nv])
        assert torch.equal(trk["visual_features_list"][o][o_off[i]::st][:L[i]].t(), vid["so_features_list"][i][nv:2 * nv])


def test_pair_table_filters_like_the_loader():
    durs = [[0, 100], [90, 200], [97, 300], [0, 3]]
    sids, oids = [0, 0, 1, 3, 0], [1, 2, 2, 0, 3]
    keep, L, s_off, o_off = MaskVRD.pair_table(durs, sids, oids, feat_stride=4, stride_offset=1, proposal_min_frames=5)
    # overlaps: 10, 3 (< proposal_min_frames), 103, 3, 3
    assert keep.tolist() == [True, False, True, False, False]
    assert L[0] == len(range(1, 10, 4)) and L[2] == len(range(1, 103, 4))
    assert (s_off[0], o_off[0]) == (90 + 1, 0 + 1) and (s_off[2], o_off[2]) == (7 + 1, 0 + 1)
