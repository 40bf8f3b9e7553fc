"""Run the UNMODIFIED reference script with its all-pairs stage replaced by the B200 engine.

The reference has no plugin API: it is one script whose stages are module-level functions that
find each other through module globals at call time, with the driver under
``if __name__ == '__main__':`` (amplicon_sorter.py:2133).  We therefore load the user's copy of
the script at run time (never vendored), execute everything except the ``__main__`` block into a
namespace, rebind ``process_list`` (:647), and execute the ``__main__`` body in that namespace.
The CLI, option semantics and every output file are the reference's own code.

    python -m amplicon_sorter_b200 --script /path/to/amplicon_sorter.py -i reads.fastq -o out -np 8
"""
from __future__ import annotations

import ast
import os
import sys


def load_reference(script_path: str):
    """-> (namespace with all top-level definitions executed, code object of the __main__ body)."""
    with open(script_path, "r") as f:
        src = f.read()
    tree = ast.parse(src, script_path)
    body, main = [], None
    for node in tree.body:
        if (isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Name)
                and node.test.left.id == "__name__"):
            main = node
        else:
            body.append(node)
    if main is None:
        raise RuntimeError(f"{script_path}: no `if __name__ == '__main__':` block found")
    ns = {"__name__": "amplicon_sorter", "__file__": script_path, "__builtins__": __builtins__}
    exec(compile(ast.Module(body=body, type_ignores=[]), script_path, "exec"), ns)
    main_code = compile(ast.Module(body=main.body, type_ignores=[]), script_path, "exec")
    return ns, main_code


def execute(ns, main_code, argv):
    """Run the reference's __main__ body with sys.argv = [script] + argv."""
    old = sys.argv
    sys.argv = [ns["__file__"]] + list(argv)
    try:
        exec(main_code, ns)
    finally:
        sys.argv = old


def install_gpu_stage(ns, device: int = 0, stats: dict | None = None, engine_factory=None):
    """Rebind process_list (:647) in the reference's namespace to the GPU implementation.

    engine_factory() -> engine lets the multi-GPU launcher hand in a dist.ShardedEngine; by default
    every call creates (and closes) a single-GPU Engine on `device`."""
    from . import host

    def process_list(self, tempfile):
        eng = engine_factory() if engine_factory else None
        return host.process_list(self, tempfile, ns["args"], engine=eng, stats_out=stats)

    process_list.__doc__ = host.process_list.__doc__
    ns["process_list"] = process_list

    # "next" row of the scope table: reads x group consensuses (:1627-1715).  One engine is kept for the
    # ~30 calls per gene group; opt out with ASB200_STAGES=process_list.
    if "process_consensuslist" in os.environ.get("ASB200_STAGES", "process_list,process_consensuslist"):
        keep = {}

        def process_consensuslist(indexes, grouplist, group_filename):
            eng = engine_factory() if engine_factory else keep.get("engine")
            if eng is None:
                from .engine import Engine

                eng = keep["engine"] = Engine(device)
            return host.process_consensuslist(indexes, grouplist, group_filename, args=ns["args"],
                                              comparelist2=ns["comparelist2"], similar=ns["similar"], engine=eng)

        process_consensuslist.__doc__ = host.process_consensuslist.__doc__
        ns["process_consensuslist"] = process_consensuslist

        # consensus x consensus (:1139-1158): intercept the worker-pool dispatch for that one worker
        if "iden_consensus" in os.environ.get("ASB200_STAGES", "iden_consensus"):
            original_do_parallel = ns["do_parallel"]

            def do_parallel(outputfolder, nprocesses, consensus_tempfile, worker, stringx, group_filename):
                if getattr(worker, "__name__", "") != "iden_consensus":
                    return original_do_parallel(outputfolder, nprocesses, consensus_tempfile, worker, stringx, group_filename)
                eng = engine_factory() if engine_factory else keep.get("engine")
                if eng is None:
                    from .engine import Engine

                    eng = keep["engine"] = Engine(device)
                return host.iden_consensus_files(outputfolder, consensus_tempfile, stringx, engine=eng)

            ns["do_parallel"] = do_parallel
    ns["check_version"] = lambda version: None  # :39-72 fetches GitHub and may sleep 10 s; not part of the path


def find_script(explicit: str | None) -> str:
    cands = [explicit, os.environ.get("AMPLICON_SORTER_PY"), os.path.join(os.getcwd(), "amplicon_sorter.py")]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise SystemExit("amplicon_sorter.py not found: pass --script PATH or set AMPLICON_SORTER_PY")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    script, device = None, int(os.environ.get("ASB200_DEVICE", "0"))
    if "--script" in argv:
        k = argv.index("--script")
        script = argv[k + 1]
        del argv[k:k + 2]
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_distributed(find_script(script), argv)
    ns, main_code = load_reference(find_script(script))
    install_gpu_stage(ns, device)
    execute(ns, main_code, argv)


def main_distributed(script, argv):
    """torchrun entry: rank 0 runs the reference, the other ranks serve its all-pairs calls."""
    from . import dist
    from .engine import Engine

    r, w, dev = dist.init_from_env()
    engine = Engine(dev.index if dev.type == "cuda" else 0)
    if r != 0:
        return dist.worker_loop(engine, dev)
    sharded = dist.ShardedEngine(engine, dev)
    try:
        ns, main_code = load_reference(script)
        install_gpu_stage(ns, engine_factory=lambda: sharded)
        execute(ns, main_code, argv)
    finally:
        sharded.close()


if __name__ == "__main__":
    main()
