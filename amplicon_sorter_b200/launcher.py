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


ALL_STAGES = ("process_list", "process_consensuslist", "iden_consensus", "groups")


def enabled_stages() -> set:
    """ASB200_STAGES = comma-separated subset of ALL_STAGES to run on the GPU (default: all of them).  Every stage is
    independent of the others; a stage that is not listed stays the reference's own CPU code."""
    raw = os.environ.get("ASB200_STAGES")
    if raw is None:
        return set(ALL_STAGES)
    want = {t.strip() for t in raw.split(",") if t.strip()}
    unknown = want - set(ALL_STAGES)
    if unknown:
        raise SystemExit(f"ASB200_STAGES: unknown stage(s) {sorted(unknown)}; known: {', '.join(ALL_STAGES)}")
    return want


def loud(fn, passthrough=()):
    """The reference skips an input file on ANY exception (`except Exception: continue`, amplicon_sorter.py:2184-2185)
    without a word.  That is its protocol for `No reads to compare` (:702-706, :768-772) -- but a GPU stage has failure
    modes the CPU path does not have (no device, out of memory, a read too long for the band, a lost rank), and those
    must not turn into silently missing output files and exit code 0.  Anything but `passthrough` raised by a replaced
    stage is therefore printed and turned into SystemExit, which the reference's handler does not catch."""
    import functools
    import traceback

    @functools.wraps(fn)
    def wrapper(*a, **k):
        try:
            return fn(*a, **k)
        except passthrough:
            raise
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            raise SystemExit(f"amplicon_sorter_b200: stage {fn.__name__} failed (see the traceback above); "
                             "stopping instead of skipping the input file silently")

    return wrapper


def install_gpu_stage(ns, device: int = 0, stats: dict | None = None, engine_factory=None, stages: set | None = None):
    """Rebind process_list (:647) -- and the "next" stages -- in the reference's namespace to the GPU implementations.

    ONE engine serves the whole run (every input file, every stage) on `device`; engine_factory() -> engine lets the
    multi-GPU launcher hand in a dist.ShardedEngine instead.  The later CPU stages of the reference fork worker
    processes (`Process(target=make_consensus)`, :1189-1198) while this engine's CUDA context is alive: the children
    never touch the library (they only run the reference's own functions), which is the supported way to fork after
    CUDA; they must be FORKED though -- the reference's functions live in a namespace that cannot be pickled for
    spawn/forkserver -- so the start method is pinned to 'fork' here, as the reference itself does with -mac (:2136)."""
    import multiprocessing

    from . import host

    try:
        multiprocessing.set_start_method("fork")
    except RuntimeError:  # already set by the embedding program
        if multiprocessing.get_start_method() != "fork":
            print("asb200: multiprocessing start method is not 'fork'; the reference's worker processes need it", file=sys.stderr)
    stages = enabled_stages() if stages is None else stages
    keep: dict = {}

    def the_engine():
        eng = engine_factory() if engine_factory else keep.get("engine")
        if eng is None:
            from .engine import Engine

            eng = keep["engine"] = Engine(device)
        return eng

    def small_stage_engine():
        """Engine for the stages that do not shard (a few milliseconds of work each: reads x consensuses, consensus x
        consensus, the consumers of the tempfile): under torchrun that is rank 0's own engine, not the sharded facade."""
        eng = the_engine()
        return getattr(eng, "engine", eng)  # dist.ShardedEngine.engine = the local engine

    ns["__asb_engine__"] = the_engine
    if "process_list" in stages:
        def process_list(self, tempfile):
            return host.process_list(self, tempfile, ns["args"], engine=the_engine(), stats_out=stats)

        process_list.__doc__ = host.process_list.__doc__
        ns["process_list"] = loud(process_list, passthrough=(host.NoReadsToCompare,))

    # "next" row of the scope table: reads x group consensuses (:1627-1715)
    if "process_consensuslist" in stages:
        def process_consensuslist(indexes, grouplist, group_filename):
            return host.process_consensuslist(indexes, grouplist, group_filename, args=ns["args"],
                                              comparelist2=ns["comparelist2"], similar=ns["similar"], engine=small_stage_engine())

        process_consensuslist.__doc__ = host.process_consensuslist.__doc__
        ns["process_consensuslist"] = loud(process_consensuslist)

    # consensus x consensus (:1139-1158): intercept the worker-pool dispatch for that one worker
    if "iden_consensus" in stages:
        original_do_parallel = ns["do_parallel"]

        def do_parallel(outputfolder, nprocesses, consensus_tempfile, worker, stringx, group_filename):
            if getattr(worker, "__name__", "") != "iden_consensus":
                return original_do_parallel(outputfolder, nprocesses, consensus_tempfile, worker, stringx, group_filename)
            return loud(host.iden_consensus_files)(outputfolder, consensus_tempfile, stringx, engine=small_stage_engine())

        ns["do_parallel"] = do_parallel
    ns["check_version"] = lambda version: None  # :39-72 fetches GitHub and may sleep 10 s; not part of the path
    # "next" rows 3-4: the consumers of the tempfile (SSG, best-hit filter, grouping)
    if "groups" in stages:
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

    ns["__asb_update_list_filter__"] = loud(update_list_filter)
    ns["__asb_read_indexes_filter__"] = loud(read_indexes_filter)
    ns["__asb_groups__"] = loud(make_groups)
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
        ns["SSG"] = loud(SSG)
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
    stages = enabled_stages()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_distributed(find_script(script), argv, stages)
    # Fail loudly BEFORE the reference runs: there is no CPU fallback, and a missing GPU or a missing libasb200.so must
    # not surface as "every input file skipped".
    from .engine import Engine

    try:
        engine = Engine(device)
    except Exception as exc:
        raise SystemExit(f"amplicon_sorter_b200: cannot start the CUDA engine on device {device}: {exc}")
    ns, main_code = load_reference(find_script(script))
    install_gpu_stage(ns, device, engine_factory=lambda: engine, stages=stages)
    try:
        execute(ns, main_code, argv)
    finally:
        engine.close()


def main_distributed(script, argv, stages=None):
    """torchrun entry: rank 0 runs the reference, the other ranks serve its all-pairs calls."""
    import torch

    from . import dist
    from .engine import Engine

    r, w, dev = dist.init_from_env()
    engine, err = None, None
    try:
        if dev.type == "cuda":
            # engine and torch.distributed share ONE non-default stream: the NCCL gather reads what the engine wrote
            stream = torch.cuda.Stream(dev)
            torch.cuda.set_stream(stream)
            engine = Engine(dev.index, stream=stream.cuda_stream)
        else:
            engine = Engine(0)
    except Exception as exc:  # noqa: BLE001 -- every rank learns about it below
        err = exc
    try:
        dist.agree(err is None, dev, "engine start-up")  # one bad GPU must not leave the other ranks parked in a broadcast
    except dist.DistributedAbort:
        if err is not None:
            print(f"amplicon_sorter_b200: rank {r}: cannot start the CUDA engine: {err}", file=sys.stderr)
        raise
    if r != 0:
        return dist.worker_loop(engine, dev)
    sharded = dist.ShardedEngine(engine, dev)
    try:
        ns, main_code = load_reference(script)
        install_gpu_stage(ns, engine_factory=lambda: sharded, stages=stages)
        execute(ns, main_code, argv)
    finally:
        # after a DistributedAbort the workers are gone (close() then skips the 'stop' broadcast); after any other error
        # on rank 0 they sit in worker_loop's broadcast, where 'stop' reaches them
        sharded.close()


if __name__ == "__main__":
    main()
