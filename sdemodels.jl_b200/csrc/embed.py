"""Embed text files as C string constants (chunked: string literals have length limits)."""
import sys

out = sys.argv[1]
with open(out, "w") as fh:
    for spec in sys.argv[2:]:
        name, path = spec.split("=")
        data = open(path, encoding="utf-8").read()
        fh.write(f"extern const char *{name};\nconst char *{name} =\n")
        for line in data.split("\n"):
            esc = line.replace("\\", "\\\\").replace('"', '\\"')
            fh.write(f'  "{esc}\\n"\n')
        fh.write(";\n")
