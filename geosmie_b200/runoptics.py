#!/usr/bin/env python3
"""Top-level caller for optics table generation (mirror of src/geosmie/runoptics.py; same options and file names).

    python -m geosmie_b200.runoptics --name geosparticles/su.json --dest out
    torchrun --nproc-per-node 8 -m geosmie_b200.runoptics --name geosparticles/ss.json     # cells sharded over 8 GPUs
"""
import os
from optparse import OptionParser

from . import dointegration
from . import particleparams as pp


# (flags, dest, default, help) -- the option set of the reference CLI (src/geosmie/runoptics.py:28-51) plus --dense
_OPTIONS = [
    (("--name",), "name", "", "Particle file to use (default=%s)" % ""),
    (("--namelist",), "namelist", "",
     "File with list of particle files (to be passed to --file) to run iteratively. If used, overrides --file (default=%s)" % ""),
    # the reference keeps --shape inside a string literal (runoptics.py:35-39: switched off); accepted here with its one
    # supported value so that scripts written against older revisions keep running
    (("--shape",), "shape", "mie", "Particle shape to use %s (default=%s)" % (['mie'], "mie")),
    (("--datatype",), "datatype", "json", "Particle data type to use %s (default=%s)" % (['json'], "json")),
    (("--dest",), "dest", ".", "Output directory to use (default=%s)" % "."),
]
_FLAGS = [
    (("-c", "--classic"), "classic", "write output filename is legacy dimensioning"),
    (("--dense",), "dense", "also evaluate particles whose size-distribution weight is exactly zero (reference-equivalent work)"),
]


def _parse(argv):
    parser = OptionParser(usage="Usage: %prog", version='0.0.1')
    for flags, dest, default, text in _OPTIONS:
        parser.add_option(*flags, dest=dest, default=default, help=text)
    for flags, dest, text in _FLAGS:
        parser.add_option(*flags, action="store_true", dest=dest, default=False, help=text)
    options, _ = parser.parse_args(argv)
    if options.shape not in ['mie']:
        parser.error("shape must be one of: %s (kernel-mode dust is selected by \"mode\" in the particle file)" % (['mie']))
    if options.datatype not in ['json']:
        parser.error("data type must be one of: %s" % (['json']))
    if not os.path.exists(options.dest):
        parser.error("Output directory (--dest) does not exist")
    if not options.name and not options.namelist:
        parser.error("non-empty particle name or namelist required (use --name particlename or --namelist namelist)")
    if options.namelist:
        if not os.path.exists(options.namelist):
            parser.error("Namelist %s does not exist" % options.namelist)
        with open(options.namelist) as fp:
            names = [line.strip() for line in fp.readlines()]
    else:
        names = [options.name]
    return parser, options, names


def main(argv=None):
    parser, options, namelist = _parse(argv)

    comm = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from . import dist
        comm = dist.Comm.from_env()
    rank = 0 if comm is None else comm.rank

    for fni, fn in enumerate(namelist):
        print("Starting particle file %s, %d of %d" % (fn, fni + 1, len(namelist)))
        if not os.path.exists(fn):
            fn1 = "%s.json" % fn     # try with added json in case it was omitted
            if not os.path.exists(fn1):
                parser.error("File %s doesn't exist" % fn)
            fn = fn1
        params = pp.getParticleParams(fn, options.datatype)
        particlename = fn.split('/')[-1].replace(".json", "")
        # Species with "hydrophobic": true get a hydrophobic bin prepended (RH-index-0 values at all RH).  The reference
        # writes the table, renames it, re-reads it in hydrophobic.doConversion and writes it again (runoptics.py:113-121);
        # here the same rule is applied to the arrays in memory and the file is written once.
        hp = bool(params.get("hydrophobic"))
        out = dointegration.fun(fn, options.datatype, options.dest, options.classic, elide=not options.dense, comm=comm,
                                write=not hp)
        opfn = "optics_%s.nomom.legacy.nc4" % particlename if options.classic else "optics_%s.nomom.nc4" % particlename
        if rank == 0 and hp:
            print("Starting hydrophobic bin handling")
            dointegration.write_table(particlename, options.dest, out, options.classic, hydrophobic_bin=True)
        print("Done, output file: %s" % opfn)
    if comm is not None:
        comm.close()


if __name__ == "__main__":
    main()
