#!/usr/bin/env python3
"""CLI of the GSF moment expansion (mirror of src/gsf/rungsf.py; same options)."""
import os
from optparse import OptionParser

from . import convertncdf


def main(argv=None):
    parser = OptionParser(usage="Usage: %prog", version='0.0.1')
    parser.add_option("--filename", dest="filename", default="", help="Optical table file to use (default=%s)" % (""))
    parser.add_option("--dest", dest="dest", default=".", help="Output directory (default=%s)" % ("."))
    parser.add_option("--mode", dest="mode", default="pygeos", help="Input file format (default=%s)" % ("pygeos"))
    parser.add_option("--rhop", dest="rhop", default=1000.0,
                      help="Particle density (not needed/used for modes pygeos, legendre) (default=%s)" % (1000.0))
    (options, args) = parser.parse_args(argv)
    if not os.path.exists(options.filename):
        parser.error("Input file path (--filename) does not exist")
    convertncdf.convertFile(options.filename, options.dest, options.mode, options.rhop)
    print('Done!')


if __name__ == "__main__":
    main()
