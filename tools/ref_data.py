"""Helpers that read the reference's ice-field pickles (when the reference tree is present).
Used only by fixture generators and by tests that are skipped when /root/reference is absent."""
import os
import pickle

REF = os.environ.get("BENCHPUSH_REFERENCE", "/root/reference")
ICE_DIR = os.path.join(REF, "benchpush/environments/ship_ice_nav/ice_environments")


def have_reference():
    return os.path.isdir(ICE_DIR)


def load_experiment(concentration):
    fn = os.path.join(ICE_DIR, "experiments_%d_100_r06_d40x12.pk" % int(concentration * 100))
    with open(fn, "rb") as f:
        d = pickle.load(f)
    return d["exp"][concentration]
