"""``input.json`` reader (reference ``cli/json_loader.py``)."""
import json


def load_input_json(input_file: str = "input.json") -> dict:
    """Settings dictionary: keys ``dim, H, R, Q, P, dt, nsteps[, smooth]``."""
    with open(input_file, "r") as fh:
        return json.load(fh)
