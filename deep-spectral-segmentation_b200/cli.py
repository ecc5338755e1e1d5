"""python-fire compatible command line (fire is not installed in this environment):

    python extract.py <command> [positional ...] --flag value --flag=value --boolflag --noboolflag

Values are parsed as Python literals when possible (fire's behaviour), otherwise kept as strings."""
from __future__ import annotations

import ast
import inspect
import sys
from typing import Callable, Dict, List


def _parse_value(text: str):
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def parse_args(fn: Callable, argv: List[str]):
    sig = inspect.signature(fn)
    names = list(sig.parameters)
    args, kwargs = [], {}
    i = 0
    while i < len(argv):
        tok = argv[i]
        if tok.startswith("--"):
            body = tok[2:]
            if "=" in body:
                key, val = body.split("=", 1)
                kwargs[key.replace("-", "_")] = _parse_value(val)
            else:
                key = body.replace("-", "_")
                nxt = argv[i + 1] if i + 1 < len(argv) else None
                if key not in names and key.startswith("no") and key[2:] in names:
                    kwargs[key[2:]] = False
                elif nxt is None or (nxt.startswith("--") and not _is_number(nxt)):
                    kwargs[key] = True
                else:
                    kwargs[key] = _parse_value(nxt)
                    i += 1
        else:
            args.append(_parse_value(tok))
        i += 1
    unknown = [k for k in kwargs if k not in names]
    if unknown:
        raise SystemExit(f"ERROR: unknown flag(s) {unknown} for {fn.__name__}; accepted: {names}")
    return args, kwargs


def _is_number(s: str) -> bool:
    try:
        float(s)
        return True
    except ValueError:
        return False


def Fire(commands: Dict[str, Callable], argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print("commands:\n  " + "\n  ".join(f"{k}{inspect.signature(v)}" for k, v in commands.items()))
        return None
    name = argv[0]
    if name not in commands:
        raise SystemExit(f"ERROR: unknown command {name!r}; available: {sorted(commands)}")
    fn = commands[name]
    if any(a in ("-h", "--help") for a in argv[1:]):
        print(f"{name}{inspect.signature(fn)}\n{inspect.getdoc(fn) or ''}")
        return None
    args, kwargs = parse_args(fn, argv[1:])
    return fn(*args, **kwargs)
