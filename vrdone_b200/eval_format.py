"""Consumer side of the forward's result dict: the per-video conversion the reference's eval loop applies right after
``model(proposal)`` (reference utils/evaluate.py:12-73 ``EvaluationFormatConvertor.to_eval_format_pr``, called from
eval.py:151), SURVEY.md section 8f row 3.

The conversion only re-labels ids and passes ``so_trajs[i][0] / [1]`` on as ``sub_traj`` / ``obj_traj``; it never looks
inside the trajectories.  With ``model.lazy_trajs = True`` those entries stay (n, 4) float32 array views until a consumer
really needs nested lists (``trajs="list"``), which is where the ~10^5 Python floats per video of the eager format go.

The id -> name tables live in the reference's ``dataloaders/category.py`` (data, not code): they are passed in, e.g.
``EvaluationFormatConvertor("vidor", vidor_category_id_to_name, vidor_pred_id_to_name)``; without tables the ids themselves
are reported (enough for tests and for consumers that map later).
"""
from __future__ import annotations

from typing import Mapping, Optional

from .maskvrd import LazyTrajs


class _Identity(dict):
    def __missing__(self, key):
        return key


class EvaluationFormatConvertor:
    """Same interface as the reference class (``to_eval_format_pr(video_name, pr_triplet)`` -> ``{video_name: [relation
    dicts]}``), with the category tables injected.  ``trajs``: "list" (reference format: nested Python lists), "array" (keep
    (n, 4) float32 numpy views when the result carries ``LazyTrajs``; lists are passed through untouched)."""

    def __init__(self, dataset_type: str, entity_id_to_name: Optional[Mapping] = None, pred_id_to_name: Optional[Mapping] = None,
                 trajs: str = "list"):
        self.dataset_type = dataset_type.lower()
        if self.dataset_type not in ("vidvrd", "vidor"):
            raise NotImplementedError(dataset_type)
        assert trajs in ("list", "array")
        self.entity_id_to_name = entity_id_to_name if entity_id_to_name is not None else _Identity()
        self.pred_id_to_name = pred_id_to_name if pred_id_to_name is not None else _Identity()
        self.trajs = trajs

    def _reset_video_name(self, video_name: str) -> str:
        if self.dataset_type == "vidor":          # e.g. "0001_3598080384" -> "3598080384" (evaluate.py:25-31)
            parts = video_name.split("_")
            assert len(parts) == 2
            return parts[1]
        return video_name                         # ImageNet-VidVRD names are used as they are

    def to_eval_format_pr(self, video_name: str, pr_triplet) -> dict:
        video_name = self._reset_video_name(video_name)
        if pr_triplet is None:                    # the reference's loop skips such videos before converting (eval.py:148-149)
            return {video_name: []}
        trajs = pr_triplet["so_trajs"]
        lazy = isinstance(trajs, LazyTrajs)
        out = []
        for i, (s_cat, p_cat, o_cat) in enumerate(pr_triplet["triplets"]):
            start, end = pr_triplet["pred_durations"][i]
            if lazy and self.trajs == "array":
                sub, obj = trajs.arrays(i)
            else:
                sub, obj = trajs[i][0], trajs[i][1]
            assert len(sub) == len(obj) == end - start
            out.append({"triplet": [self.entity_id_to_name[s_cat], self.pred_id_to_name[p_cat], self.entity_id_to_name[o_cat]],
                        "duration": (start, end), "score": float(pr_triplet["triple_scores_avg"][i]),
                        "sub_traj": sub, "obj_traj": obj})
        return {video_name: out}
