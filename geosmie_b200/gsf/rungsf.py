#!/usr/bin/env python3
"""CLI of the GSF moment expansion (mirror of src/gsf/rungsf.py; same options)."""
import os
from optparse import OptionParser

from . import convertncdf


# (flag, default, help) -- the option set of the reference CLI (src/gsf/rungsf.py:12-26)
_OPTIONS = [
    ("filename", "", "Optical table file to use (default=%s)" % ""),
    ("dest", ".", "Output directory (default=%s)" % "."),
    ("mode", "pygeos", "Input file format (default=%s)" % "pygeos"),
    ("rhop", 1000.0, "Particle density (not needed/used for modes pygeos, legendre) (default=%s)" % 1000.0),
]


def main(argv=None):
    parser = OptionParser(usage="Usage: %prog", version='0.0.1')
    for name, default, text in _OPTIONS:
        parser.add_option("--" + name, dest=name, default=default, help=text)
    options, _ = parser.parse_args(argv)
    if not os.path.exists(options.filename):
        parser.error("Input file path (--filename) does not exist")
    convertncdf.convertFile(options.filename, options.dest, options.mode, options.rhop)
    print('Done!')


if __name__ == "__main__":
    main()
