"""Loader for the committed case fixtures under tests/golden/<case>/ (copies of the reference's own *input* files
for BASELINE configs 1-3, made by tests/golden/make_fixtures.py; /root/reference does not exist on the GPU box)."""
import os


def load(case_mod, path, scheme=None, control=None, flow=None):
    blocks = case_mod.load_case(path, scheme_override=scheme, control_override=control, flow_override=flow)
    restart = os.path.join(path, "restart")
    if os.path.isdir(restart):
        for b, blk in enumerate(blocks):
            case_mod.read_tecplot_state(os.path.join(restart, "process_%02d.dat" % b), blk)
    return blocks
