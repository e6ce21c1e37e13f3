"""Success / Precision of a tracker over a set of tracklets — what the reference's Lightning `test_step` /
`validation_step` accumulate (models/base_model.py:88-117: every frame's overlap and centre distance go into
TorchSuccess / TorchPrecision, whose `compute()` is the area under the curve)."""
from ..utils.metrics import Precision, Success


def evaluate(model, sequences, progress=None):
    """sequences: iterable of tracklets (lists of {"pc", "3d_bbox"}).  Returns {"success", "precision", "frames", "results"}."""
    succ, prec, results, frames = Success(), Precision(), [], 0
    for i, seq in enumerate(sequences):
        ious, dists, boxes = model.evaluate_one_sequence(seq)
        succ(ious)
        prec(dists)
        results.append(boxes)
        frames += len(seq)
        if progress is not None:
            progress(i, succ.compute(), prec.compute())
    return {"success": succ.compute(), "precision": prec.compute(), "frames": frames, "results": results}
