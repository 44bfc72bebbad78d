"""Command line: ``python -m xpore_b200.cli diffmod --config cfg.yml [...]``.

Same sub-command and flags as the reference's ``xpore diffmod`` (``xpore/scripts/xpore.py:40-52``):
``--config`` (required), ``--n_processes``, ``--save_models``, ``--resume``, ``--ids``.
``postprocessing --diffmod_dir`` (``xpore.py:54-59``) is the majority-direction filter of
``xpore_b200.postprocessing``.  ``dataprep`` is outside this repository's scope.
"""
import argparse
import sys


def parse_options(argv):
    parser = argparse.ArgumentParser(prog="xpore", description="xpore_b200 - B200-native xPore diffmod")
    sub = parser.add_subparsers(dest="cmd")
    sub.required = True
    p = sub.add_parser("diffmod", help="run mode for performing differential modification analysis")
    req = p.add_argument_group("required arguments")
    req.add_argument("--config", dest="config", help="YAML configuraion filepath.", required=True)
    p.add_argument("--n_processes", dest="n_processes", type=int, default=1,
                   help="number of host processes that parse data.json and select positions (the fit runs on the GPU).")
    p.add_argument("--save_models", dest="save_models", default=False, action="store_true",
                   help="with this argument, the program will save the model parameters for each id "
                        "(<out>/models/<id>.npz, and <id>.hdf5 in the reference's layout when h5py is installed).")
    p.add_argument("--resume", dest="resume", default=False, action="store_true",
                   help="with this argument, the program will resume from the previous run.")
    p.add_argument("--ids", dest="ids", default=[], nargs="*", help="gene or transcript ids to model.")
    p.add_argument("--batch_genes", dest="batch_genes", type=int, default=8192,
                   help="at most this many ids are ingested and fitted per GPU call.")
    p.add_argument("--batch_bytes", dest="batch_bytes", type=int, default=256 << 20,
                   help="... or this many bytes of data.json text, whichever comes first.")
    p.add_argument("--csr_cache", dest="csr_cache", default=None,
                   help="binary CSR cache of the inputs (xpore_b200.csrcache): read instead of data.json when it "
                        "matches the inputs, written on first use otherwise.")
    p.add_argument("--timing", dest="timing", default=None,
                   help="write the wall-clock seconds of the pipeline stages (ingest / GPU call / rows / write) to this JSON file.")
    q = sub.add_parser("postprocessing", help="run mode for post-processing diffmod.table, the result table from "
                                              "differential modification analysis.")
    qreq = q.add_argument_group("required arguments")
    qreq.add_argument("--diffmod_dir", dest="diffmod_dir", required=True,
                      help="diffmod directory path, the output from xpore-diffmod.")
    return parser.parse_args(argv[1:])


def main(argv=None):
    argv = sys.argv if argv is None else argv
    options = parse_options(argv)
    if options.cmd == "diffmod":
        import os
        from . import _lib
        # the CUDA context comes up on a background thread from here on, under the imports and the index reads
        _lib.warm_up_async(int(os.environ.get("LOCAL_RANK", "0")))
        from .diffmod import diffmod
        diffmod(options)
    elif options.cmd == "postprocessing":
        from .postprocessing import postprocessing
        postprocessing(options)


if __name__ == "__main__":
    main(sys.argv)
    # Single-process runs leave without tearing down: the outputs are closed, and unwinding the CUDA context and some
    # hundred MB of page-locked staging buffers costs 0.3 - 0.8 s of wall clock that the operating system does for free.
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) == 1:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
