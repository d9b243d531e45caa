"""Command line flags of ``track_estimator`` (same flags and destinations as the reference
``cli/argument_parser.py:7-93``)."""
import argparse

from .. import __version__

_FLAGS = [
    (("-i", "--input"), dict(dest="input_file", default="input.json", help="Filepath to the input JSON file")),
    (("-o", "--output"), dict(dest="output_prefix", default="output", help="Output file prefix")),
    (("-t", "--track-file"), dict(dest="track_file", required=True, help="Filepath to the ship track data")),
    (("-s", "--ship-id"), dict(dest="ship_id", required=True, help="Ship ID")),
    (("-lat", "--latitude-id"), dict(dest="lat_id", required=True, help="Name of the latitude column")),
    (("-lon", "--longitude-id"), dict(dest="lon_id", required=True, help="Name of the longitude column")),
    (("-ic", "--id-col"), dict(dest="id_col", required=True, help="Name of the ship ID column")),
    (("-rts", "--rts-smoother"), dict(dest="apply_rts_smoother", action="store_true",
                                     help="Apply the Rauch-Tung-Striebel (RTS) smoother")),
    (("-rev", "--reverse"), dict(dest="reverse", action="store_true", help="Reverse the trajectory")),
]


def create_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(
        description=f"Ship track estimator {__version__} command line interface",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
    )
    for names, kw in _FLAGS:
        parser.add_argument(*names, **kw)
    parser.add_argument("-v", "--version", action="version", version="%(prog)s {version}".format(version=__version__))
    return parser
