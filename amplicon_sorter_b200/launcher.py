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
    ns["__asb_functions__"] = {n.name: n for n in body if isinstance(n, ast.FunctionDef)}  # for install_group_stage
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
    keep: dict = {}  # one engine for all the small stages of a run

    def small_stage_engine():
        """Engine for the stages that do not shard (a few milliseconds of work each: reads x consensuses, consensus x
        consensus, the consumers of the tempfile): under torchrun that is rank 0's own engine, not the sharded facade."""
        eng = engine_factory() if engine_factory else keep.get("engine")
        if eng is None:
            from .engine import Engine

            eng = keep["engine"] = Engine(device)
        return getattr(eng, "engine", eng)  # dist.ShardedEngine.engine = the local engine

    # "next" row of the scope table: reads x group consensuses (:1627-1715).  One engine is kept for the
    # ~30 calls per gene group; opt out with ASB200_STAGES=process_list.
    if "process_consensuslist" in os.environ.get("ASB200_STAGES", "process_list,process_consensuslist"):
        def process_consensuslist(indexes, grouplist, group_filename):
            eng = small_stage_engine()
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
                return host.iden_consensus_files(outputfolder, consensus_tempfile, stringx, engine=small_stage_engine())

            ns["do_parallel"] = do_parallel
    ns["check_version"] = lambda version: None  # :39-72 fetches GitHub and may sleep 10 s; not part of the path
    # "next" rows 3-4: the consumers of the tempfile (SSG, best-hit filter, grouping); opt out with ASB200_STAGES
    if "groups" in os.environ.get("ASB200_STAGES", "groups"):
        install_group_stage(ns, small_stage_engine, stats)


def _mentions(node, name: str) -> bool:
    return any(isinstance(n, ast.Name) and n.id == name for n in ast.walk(node))


def _is_tempfile_scan(node) -> bool:
    """`try: with open(os.path.join(outputfolder, tempfile), 'r') as tf: for line in tf: ...` (:985-1015, :1363-1393)."""
    if not isinstance(node, ast.Try) or not node.body or not isinstance(node.body[0], ast.With):
        return False
    ctx = node.body[0].items[0].context_expr
    return isinstance(ctx, ast.Call) and isinstance(ctx.func, ast.Name) and ctx.func.id == "open" and _mentions(ctx, "tempfile")


def _is_greedy_loop(node) -> bool:
    """`for x in templist: for s in grouplist: ... else: grouplist.append({...})` (:1022-1031, :1403-1409)."""
    return (isinstance(node, ast.For) and isinstance(node.iter, ast.Name) and node.iter.id == "templist" and len(node.body) >= 1
            and any(isinstance(b, ast.For) and isinstance(b.iter, ast.Name) and b.iter.id == "grouplist" and b.orelse for b in node.body))


def _is_merge_call(node) -> bool:
    return (isinstance(node, ast.Assign) and isinstance(node.value, ast.Call) and isinstance(node.value.func, ast.Name)
            and node.value.func.id == "merge_groups" and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name)
            and node.targets[0].id == "grouplist")


def _rewrite_consumer(fn: ast.FunctionDef, filter_call: str, update_with_list: bool = False):
    """Copy of `fn` whose scan of the tempfile and whose greedy-grouping + merge_groups statements are replaced by
    calls into this package; every other statement stays the reference's own.  None if the shapes are not found."""
    import copy

    fn = copy.deepcopy(fn)
    done = {"scan": 0, "greedy": 0, "merge": 0}

    def visit(stmts):
        out = []
        for st in stmts:
            if _is_tempfile_scan(st):
                out.extend(ast.parse(filter_call).body)
                done["scan"] += 1
                continue
            if _is_greedy_loop(st):
                out.extend(ast.parse(f"grouplist = __asb_groups__(templist, {bool(update_with_list)})").body)
                done["greedy"] += 1
                continue
            if _is_merge_call(st) and done["greedy"] == 1 and done["merge"] == 0:
                done["merge"] += 1  # merge_groups' fixed point is what __asb_groups__ returns (and it prints its two lines)
                continue
            for field in ("body", "orelse", "finalbody"):
                sub = getattr(st, field, None)
                if isinstance(sub, list) and sub and isinstance(sub[0], ast.stmt):
                    setattr(st, field, visit(sub))
            if isinstance(st, ast.Try):
                for h in st.handlers:
                    h.body = visit(h.body)
            out.append(st)
        return out

    fn.body = visit(fn.body)
    if done != {"scan": 1, "greedy": 1, "merge": 1}:
        return None
    return ast.fix_missing_locations(fn)


def install_group_stage(ns, engine_getter, stats: dict | None = None):
    """ "Next" rows 3-4 of the scope table: SSG (:809-835), and inside update_list (:966-1055) and read_indexes
    (:1340-1460) the scan of the tempfile (best-hit filter) and greedy grouping + merge_groups, on the GPU.

    The reference offers no seam inside those two functions, so the user's own copy of each is re-compiled with
    exactly three statements swapped (found by shape; if the script does not have them, the function is left alone)."""
    from . import groups

    fns = ns.get("__asb_functions__", {})

    def SSG(tempfile):
        print("Estimating the ssg value for this dataset")  # :812
        lines = groups.lines_for(os.path.join(ns["args"].outputfolder, tempfile))
        est = groups.ssg_estimate(engine_getter(), lines, stats)
        if est is not None:
            print("-> Estimated ssg = " + str(est))  # :834
        return est

    def update_list_filter(path):
        try:
            lines = groups.lines_for(path)
        except FileNotFoundError:
            sys.exit()  # :1014-1015
        templist, *_ = groups.best_hits(engine_getter(), lines, stats=stats)
        return templist

    def read_indexes_filter(path, similar_species_groups, indexes):
        try:
            lines = groups.lines_for(path)
        except FileNotFoundError:
            return []  # :1392-1393
        templist, *_ = groups.best_hits(engine_getter(), lines, similar_species_groups, indexes, stats=stats)
        return templist

    def make_groups(templist, update_with_list):
        n_greedy, grouplist = groups.make_groups(engine_getter(), templist, update_with_list, stats)
        if n_greedy > 1:  # merge_groups' two progress lines (:1061, :1084)
            print("--> Number of groups before merge: " + str(n_greedy))
            print("--> Number of groups after merge: " + str(len(grouplist)))
        return grouplist

    ns["__asb_update_list_filter__"] = update_list_filter
    ns["__asb_read_indexes_filter__"] = read_indexes_filter
    ns["__asb_groups__"] = make_groups
    calls = {"update_list": "templist = __asb_update_list_filter__(os.path.join(outputfolder, tempfile))",
             "read_indexes": "templist = __asb_read_indexes_filter__(os.path.join(outputfolder, tempfile), similar_species_groups, indexes)"}
    installed = []
    for name, call in calls.items():
        new = _rewrite_consumer(fns[name], call, update_with_list=(name == "read_indexes")) if name in fns else None
        if new is None:
            print(f"asb200: {name} does not have the expected shape; its tempfile scan stays on the CPU", file=sys.stderr)
            continue
        exec(compile(ast.Module(body=[new], type_ignores=[]), ns.get("__file__", "<reference>"), "exec"), ns)
        installed.append(name)
    if "SSG" in ns:
        ns["SSG"] = SSG
        installed.append("SSG")
    return installed


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
    # Fail loudly BEFORE the reference runs: its per-file `except Exception: continue` (:2184) would turn a missing GPU or
    # a missing libasb200.so into a silently skipped input file.  There is no CPU fallback.
    from .engine import Engine

    try:
        Engine(device).close()
    except Exception as exc:
        raise SystemExit(f"amplicon_sorter_b200: cannot start the CUDA engine on device {device}: {exc}")
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
