#!/usr/bin/env python
"""oracle/make_golden_shannon.py -- TEST INFRASTRUCTURE.  Dumps the reference's Shannon-limit lookup ebno(rate)
(simulation/simulation.c++:56-70 over the tables :21-52) into tests/golden/shannon.json.

The function is `static` in simulation.c++, so a scratch translation unit includes that file textually and calls it;
the four g++ portability patches of oracle/build_ref.sh are applied to a scratch copy of the reference first.  Runs
only where /root/reference exists; the JSON is committed.  Recorded per rate: the reference's limit and the first
Eb/N0 of its sweep (simulation.c++:105-106) for steps 0.5 and 0.1, the latter only where the limit is > -0.5 dB
(the reference converts a negative quotient to size_t there, which is undefined behaviour).
"""
import json
import os
import re
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CCREF_SRC", "/root/reference/src")

PROG = r'''
#include <iostream>
#include <sstream>
#include <cstdio>
#include <cstdlib>
#include "simulation/simulation.c++"
int main(int argc, char **argv) {
  for (int i = 1; i < argc; ++i) {
    const double rate = std::strtod(argv[i], nullptr);
    std::printf("%.17g %.17g\n", rate, ebno(rate));
  }
  return 0;
}
'''


def main():
    scratch = tempfile.mkdtemp()
    try:
        src = os.path.join(scratch, "src")
        shutil.copytree(REF, src)
        subprocess.check_call(["chmod", "-R", "u+w", src])
        # the portability patches of build_ref.sh (center.h, galois.h, polynomial.h)
        script = open(os.path.join(HERE, "build_ref.sh")).read()
        patch = re.search(r"python3 - \"\$SCRATCH/src\" <<'EOF'\n(.*?)\nEOF", script, re.S).group(1)
        subprocess.run(["python3", "-", src], input=patch, text=True, check=True)
        with open(os.path.join(scratch, "prog.cc"), "w") as f:
            f.write(PROG)
        exe = os.path.join(scratch, "prog")
        subprocess.check_call(["g++", "-std=c++17", "-fpermissive", "-w", "-O1", "-I" + src, "-o", exe,
                               os.path.join(scratch, "prog.cc"), os.path.join(src, "codes", "codes.c++"), "-pthread"])
        with open(os.path.join(HERE, "..", "tests", "golden", "catalogue.json")) as f:
            cat = json.load(f)
        rates = {name: e["rate"] for name, e in cat.items()}
        grid = [i / 2000.0 for i in range(1, 2000)] + [0.8, 0.800001, 0.807, 0.999, 0.9995, 0.99999]
        args = [repr(r) for r in list(rates.values()) + grid]
        out = subprocess.check_output([exe] + args, text=True).split("\n")
        rows = []
        for line in out:
            if not line.strip():
                continue
            rate, limit = (float(x) for x in line.split())
            row = {"rate": rate, "limit": limit}
            if limit > -0.5:
                for step in (0.5, 0.1):
                    tmp = int(max(limit, 0.0) / step)  # size_t conversion: truncation (limit in (-0.5, 0) truncates to 0 as well)
                    row["start_%g" % step] = (tmp + (1.0 / step)) * step
            rows.append(row)
        names = list(rates.keys())
        doc = {"source": "simulation/simulation.c++:21-70,105-106 of the reference, dumped by oracle/make_golden_shannon.py",
               "catalogue": {names[i]: rows[i] for i in range(len(names))}, "grid": rows[len(names):]}
        with open(os.path.join(HERE, "..", "tests", "golden", "shannon.json"), "w") as f:
            json.dump(doc, f, indent=0)
        print("wrote tests/golden/shannon.json:", len(rows), "rates")
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
