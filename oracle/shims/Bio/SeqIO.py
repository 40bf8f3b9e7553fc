"""Bio.SeqIO surface used by amplicon_sorter.py: parse (AS:544), index (AS:1497-1508), write (AS:1531)."""


class _Record:
    __slots__ = ("id", "seq", "qual", "description")

    def __init__(self, id_, seq, qual=None, description=""):
        self.id, self.seq, self.qual, self.description = id_, seq, qual, description

    def format(self, fmt):
        if fmt == "fastq":
            return f"@{self.description or self.id}\n{self.seq}\n+\n{self.qual or 'I' * len(self.seq)}\n"
        return f">{self.description or self.id}\n{self.seq}\n"


def parse(handle, fmt):
    if isinstance(handle, str):
        handle = open(handle, "rt")
    if fmt == "fastq":
        while True:
            h = handle.readline()
            if not h:
                return
            s = handle.readline().rstrip("\n")
            handle.readline()
            q = handle.readline().rstrip("\n")
            desc = h[1:].rstrip("\n")
            yield _Record(desc.split()[0] if desc.split() else "", s, q, desc)
    elif fmt == "fasta":
        name, chunks = None, []
        for line in handle:
            if line.startswith(">"):
                if name is not None:
                    yield _Record(name.split()[0] if name.split() else "", "".join(chunks), None, name)
                name, chunks = line[1:].rstrip("\n"), []
            else:
                chunks.append(line.strip())
        if name is not None:
            yield _Record(name.split()[0] if name.split() else "", "".join(chunks), None, name)
    else:
        raise ValueError(fmt)


def index(path, fmt):
    with open(path, "rt") as f:
        return {r.id: r for r in parse(f, fmt)}


def write(records, handle, fmt):
    if isinstance(records, _Record):
        records = [records]
    n = 0
    for r in records:
        handle.write(r.format(fmt))
        n += 1
    return n
