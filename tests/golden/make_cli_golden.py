"""Writes tests/golden/cli_flags.json from the reference's own CLI source (run in the build container, where /root/reference exists):
every `parser.add_argument(...)` of /root/reference/src/inference.py:31-89 as (flags, type, default, required, choices, action), read
from the AST (the module itself cannot be imported here: diffusers / accelerate are not installed), plus the save-path rule and the
prompt template of the loop body as literal known answers.

    python tests/golden/make_cli_golden.py
"""
import ast
import json
import os

SRC = "/root/reference/src/inference.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli_flags.json")


def main():
    tree = ast.parse(open(SRC).read())
    flags = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            names = [ast.literal_eval(a) for a in node.args]
            kw = {}
            for k in node.keywords:
                if k.arg == "help":
                    continue
                kw[k.arg] = k.value.id if isinstance(k.value, ast.Name) else ast.literal_eval(k.value)
            flags.append(dict(flags=names, line=node.lineno, **kw))
    flags.sort(key=lambda f: f["line"])
    # literal known answers of the loop body (src/inference.py:279-286, :314-324)
    category_text = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "category_text":
            category_text = ast.literal_eval(node.value)
    json.dump(dict(source="src/inference.py", flags=flags, category_text=category_text), open(OUT, "w"), indent=1)
    print(f"{len(flags)} flags -> {OUT}")


if __name__ == "__main__":
    main()
