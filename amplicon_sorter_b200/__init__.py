"""B200-native all-pairs read-similarity engine behind amplicon_sorter's process_list stage."""
__all__ = ["Engine", "process_list"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "process_list":
        from .host import process_list
        return process_list
    raise AttributeError(name)
