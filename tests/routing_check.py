"""Shared by the CPU (hostemu) and GPU routing tests: the per-chromosome routing of the reference's ByChr host restated on SAM text
(reference src/mainParallelByChromosome.c:1395-1457) and the comparison of a driver's -R output with it."""
import subprocess


def contig_names(prefix):
    return [l.split()[1].encode() for l in open(prefix + ".ann").read().split("\n")[1::2] if l]


def route_like_the_reference(sam, names, with_disc):
    """RNAME '*' -> unmapped (last destination), else its contig; without fixmate a mapped line whose RNEXT names another contig is
    also copied into 'discordant' (destination n_seqs)"""
    n = len(names)
    at = {nm: i for i, nm in enumerate(names)}
    dest = [[] for _ in range(n + 2)]
    for line in sam.split(b"\n")[:-1]:
        f = line.split(b"\t")
        chr_ = n + 1 if f[2] == b"*" else at[f[2]]
        mchr = chr_ if f[6] == b"=" else (-1 if f[6] == b"*" else at[f[6]])
        dest[chr_].append(line + b"\n")
        if with_disc and chr_ < n and 0 <= mchr < n and mchr != chr_:
            dest[n].append(line + b"\n")
    return [b"".join(d) for d in dest]


def make_chimeric(fq2_path, every=5):
    """second mates of every `every`-th pair swap their sequence and qualities with the pair 7 records on (names stay)"""
    rec = open(fq2_path, "rb").read().split(b"\n")
    n = len(rec) // 4
    for i in range(0, n - 7, every):
        j = i + 7
        rec[4 * i + 1], rec[4 * j + 1] = rec[4 * j + 1], rec[4 * i + 1]
        rec[4 * i + 3], rec[4 * j + 3] = rec[4 * j + 3], rec[4 * i + 3]
    open(fq2_path, "wb").write(b"\n".join(rec))


def check_driver(drv, args, prefix, env=None):
    """runs `drv -F` plain, with the line table (-R 1) and grouped (-R 2, -R 6); returns the bytes routed to 'discordant'"""
    names = contig_names(prefix)
    run = lambda extra: subprocess.run([drv, "-F"] + extra + args, capture_output=True, check=True, env=env).stdout
    plain = run([])
    assert plain.count(b"\n") > 1000
    # line table: every line once, in order, with the contigs the text names
    text = b""
    for chunk in run(["-R", "1"]).split(b"@@CHUNK\t")[1:]:
        size, rest = chunk.split(b"\n", 1)
        body, table = rest[:int(size)], rest[int(size):]
        pos = 0
        for row, line in zip(table.split(b"\n")[:-1], body.split(b"\n")[:-1]):
            tag, off, ln, rid, mrid, _ = row.split(b"\t")
            assert tag == b"@@LINE" and int(off) == pos and int(ln) == len(line) + 1
            f = line.split(b"\t")
            assert (b"*" if int(rid) < 0 else names[int(rid)]) == f[2]
            assert (b"*" if int(mrid) < 0 else (b"=" if mrid == rid else names[int(mrid)])) == f[6]
            pos += len(line) + 1
        assert pos == len(body) and table.count(b"\n") == body.count(b"\n")
        text += body
    assert text == plain
    # grouped text, without and with the discordant copies (the chunks partition the reads in order, so routing the whole plain
    # SAM gives the concatenation of the per-chunk routings)
    n_disc = 0
    for flags, with_disc in ((2, False), (6, True)):
        got = [b""] * (len(names) + 2)
        for chunk in run(["-R", str(flags)]).split(b"@@CHUNK\t")[1:]:
            rest = chunk.split(b"\n", 1)[1]
            while rest:
                head, rest = rest.split(b"\n", 1)
                tag, d, size = head.split(b"\t")
                assert tag == b"@@DEST"
                got[int(d)] += rest[:int(size)]
                rest = rest[int(size):]
        assert got == route_like_the_reference(plain, names, with_disc)
        n_disc = len(got[len(names)])
    return n_disc
