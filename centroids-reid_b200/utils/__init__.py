"""Drop-ins for the reference's utils/reid_metric.py and utils/eval_reid.py."""
