"""CPU: bench.py's stdout discipline -- exactly one JSON line reaches stdout whatever libraries print (NCCL writes its version
banner to stdout on multi-GPU runs; the driver parses stdout)."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_only_the_json_line_reaches_stdout():
    code = textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        import bench
        bench.claim_stdout()
        os.write(1, b"NCCL version 2.28.9+cuda12.9\\n")      # what NCCL does from C
        print("library chatter")
        bench.emit({"metric": "rays/sec", "value": 1.0})
    ''' % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0])["metric"] == "rays/sec"
    assert "NCCL version" in r.stderr and "library chatter" in r.stderr
