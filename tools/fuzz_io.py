"""Development tool: mutate the format fixtures of tests/golden/io (byte flips, truncation,
insertions, huge sizes) and load them. Meant to be run against a sanitizer build of the host
code (-fsanitize=address,undefined on tb_host.cpp / tb_io.cpp, TURTLE_B200_LIB pointing at
it): every file must either load or be refused through the error handler."""
import os, sys, random, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turtle_b200 as tb
IO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "io")
out = "/tmp/fuzz_io"; os.makedirs(out, exist_ok=True)
random.seed(7)
n_ok = n_err = 0
seeds = [(name, open(os.path.join(IO, name), "rb").read()) for name in sorted(os.listdir(IO))]
try:  # Adam7 files: written by the encoder of the tests (no fixture on disk)
    from tests.test_io_foreign_files import write_interlaced_png, terrain
    for nx, ny in ((53, 37), (3, 2), (1, 6)):
        write_interlaced_png(os.path.join(out, "seed.png"), terrain(nx, ny, 1), None)
        seeds.append(("adam7_%dx%d.png" % (nx, ny), open(os.path.join(out, "seed.png"), "rb").read()))
except ImportError:
    pass
for name, data in seeds:
    for trial in range(400):
        b = bytearray(data)
        mode = trial % 4
        if mode == 0:
            for _ in range(random.randint(1, 6)):
                b[random.randrange(len(b))] = random.randrange(256)
        elif mode == 1:
            b = b[:random.randrange(1, len(b))]
        elif mode == 2:
            i = random.randrange(len(b)); b[i:i] = bytes(random.randrange(256) for _ in range(random.randint(1, 40)))
        else:
            i = random.randrange(len(b) - 4); b[i:i + 4] = b"\xff\xff\xff\x7f"
        path = os.path.join(out, "f_%d_%s" % (trial, name))
        open(path, "wb").write(bytes(b))
        try:
            m = tb.Map(path=path); info, _ = m.meta()
            if info.nx > 0 and info.ny > 0 and info.nx * info.ny < 10**7:
                m.node(info.nx - 1, info.ny - 1)
            n_ok += 1
        except tb.TurtleError:
            n_err += 1
        os.remove(path)
print("loaded", n_ok, "refused", n_err)
