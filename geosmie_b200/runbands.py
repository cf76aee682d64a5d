#!/usr/bin/env python3
"""CLI of the band averaging (mirror of src/geosmie/runbands.py; same options)."""
import os
from optparse import OptionParser

from . import bandaverage


MODES = ['GEOS5', 'RRTMG', 'RRTMGP', 'PURDUE']
# (flag, default, help) -- the option set of the reference CLI (src/geosmie/runbands.py:14-40)
_OPTIONS = [
    ("filename", "", "Optical table file to use (default=%s)" % ""),
    ("namelist", "", "File with list of particle optics files (to be passed to --filename) to run iteratively. "
                     "If used, overrides --filename (default=%s)" % ""),
    ("partname", "", "Particle name to use in the qname variable of the output table (default=%s)" % ""),
    ("dest", ".", "Output directory (default=%s)" % "."),
    ("bandmode", "RRTMG", "Band averaging type to use %s (default=%s)" % (MODES, "RRTMG")),
    ("noIR", False, "Use the noIR option for GEOS5 band type (default=%s)" % False),
    ("useSolar", False, "Use the useSolar option (default=%s)" % False),
]


def main(argv=None):
    parser = OptionParser(usage="Usage: %prog", version='0.0.1')
    for name, default, text in _OPTIONS:
        parser.add_option("--" + name, dest=name, default=default, help=text)
    options, _ = parser.parse_args(argv)
    mode = options.bandmode.upper()
    if mode not in MODES:
        parser.error("Band type must be one of: %s" % (MODES))
    files = [options.filename]
    if options.namelist:
        if not os.path.exists(options.namelist):
            parser.error("Namelist %s does not exist" % options.namelist)
        with open(options.namelist) as fp:
            files = [line.strip() for line in fp.readlines()]
    for i, fn in enumerate(files):
        print("Starting optics file %s, %d of %d" % (fn, i + 1, len(files)))
        if not os.path.exists(fn):
            parser.error("Input file path (--filename) does not exist: %s" % fn)
        bandaverage.processFileForBandMode(fn, options.partname, options.dest, mode, options.useSolar, options.noIR)
    print('Done!')


if __name__ == "__main__":
    main()
