# SPDX-License-Identifier: Apache-2.0
"""Transcribes the end-to-end upsert scenarios of the reference (G9) into a JSON fixture.

    python tests/golden/make_upsert_goldens.py /root/reference > tests/golden/reference_upsert_scenarios.json

Source: rust/lakesoul-datafusion/src/tests/upsert_tests.rs:241-1961. Every `async fn test_*` there creates a table
(`init_table`: schema, primary keys, range partitions, first upsert), applies more upserts (`execute_upsert`, and the
batch handed to `check_upsert`) and asserts the table's content (`assert_batches_eq` on the printed table, after an
optional filter and a column selection). All inputs and expected outputs are literals in the Rust source; this script
reads them with a small translator for the handful of constructors the file uses (it needs the reference checkout,
which exists only where the fixture is produced -- the JSON is what the tests read).

Fixture format, per scenario: schema (name, type), primary_keys, range_partitions, upserts (list of column -> values),
selected columns, filter (or null), expected rows (list of column -> values, in the order the reference prints them).
"""
import datetime
import json
import re
import sys


def split_top(s: str):
    """split at top-level commas"""
    out, depth, cur, instr = [], 0, [], False
    i = 0
    while i < len(s):
        ch = s[i]
        if instr:
            cur.append(ch)
            if ch == "\\":
                cur.append(s[i + 1]); i += 1
            elif ch == '"':
                instr = False
        elif ch == '"':
            instr = True; cur.append(ch)
        elif ch in "([{":
            depth += 1; cur.append(ch)
        elif ch in ")]}":
            depth -= 1; cur.append(ch)
        elif ch == "," and depth == 0:
            out.append("".join(cur).strip()); cur = []
        else:
            cur.append(ch)
        i += 1
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


def call_args(text: str, start: int):
    """text[start] is '(' -> (argument text, index after the closing parenthesis)"""
    depth, i, instr = 0, start, False
    while i < len(text):
        ch = text[i]
        if instr:
            if ch == "\\":
                i += 1
            elif ch == '"':
                instr = False
        elif ch == '"':
            instr = True
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                return text[start + 1:i], i + 1
        i += 1
    raise ValueError("unbalanced")


class Batch(dict):
    pass


def _mk(names, values, ts=None):
    b = Batch()
    for n, v in zip(names, values):
        b[n] = list(v)
    if ts is not None:
        b[names[len(values)]] = [{"ts_us": t} for t in ts]
    return b


ENV_FUNCS = {
    "create_batch_i32": lambda names, values: _mk(names, values),
    "create_batch_optional_i32": lambda names, values: _mk(names, values),
    "create_batch_string": lambda names, values: _mk(names, values),
    "create_batch_str_or_i32": lambda names, values: _mk(names, values),
    "create_batch_i32_and_timestamp": lambda names, values, ts: _mk(names, values, ts),
}


def to_python(expr: str) -> str:
    e = expr
    e = re.sub(r"StrOrI32::V1\(([^()]*)\)", r"\1", e)
    e = re.sub(r"StrOrI32::V2\(([^()]*)\)", r"\1", e)
    e = re.sub(r"\bSome\(([^()]*)\)", r"\1", e)
    e = e.replace("vec![", "[").replace("&[", "[").replace(".to_string()", "")
    return e


def eval_expr(expr: str, lets: dict):
    expr = expr.strip()
    if expr in lets:
        v = lets[expr]
        if isinstance(v, str):
            v = eval_expr(v, lets)
            lets[expr] = v
        return v
    m = re.match(r"RecordBatch::new_empty\(", expr)
    if m:
        names = re.findall(r'"([^"]+)"', expr)
        return Batch({n: [] for n in names})
    m = re.match(r"NaiveDate::from_ymd_opt\((\d+),\s*(\d+),\s*(\d+)\)\s*\.unwrap\(\)\s*\.and_hms_micro_opt\((\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", expr)
    if m:
        y, mo, d, h, mi, s, us = map(int, m.groups())
        return datetime.datetime(y, mo, d, h, mi, s, us, tzinfo=datetime.timezone.utc)
    m = re.match(r"(\w+)\.and_utc\(\)\.timestamp_micros\(\)", expr)
    if m:
        dt = eval_expr(m.group(1), lets)
        epoch = datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)
        delta = dt - epoch
        return delta.days * 86_400_000_000 + delta.seconds * 1_000_000 + delta.microseconds
    env = dict(ENV_FUNCS)
    env["None"] = None
    for k in list(lets):
        if re.search(r"\b" + re.escape(k) + r"\b", expr):
            env[k] = eval_expr(k, lets)
    return eval(to_python(expr), {"__builtins__": {}}, env)  # noqa: S307 (restricted grammar, trusted fixture source)


def parse_table(lines):
    rows = [ln for ln in lines if ln.startswith("|")]
    header = [c.strip() for c in rows[0].strip("|").split("|")]
    out = []
    for r in rows[1:]:
        cells = [c.strip() for c in r.strip("|").split("|")]
        out.append(dict(zip(header, cells)))
    return header, out


def main(ref_root: str):
    path = ref_root.rstrip("/") + "/rust/lakesoul-datafusion/src/tests/upsert_tests.rs"
    src = "\n".join(re.sub(r"^(\s*)//.*$", r"\1", ln) for ln in open(path).read().split("\n"))  # full-line comments
    line_of = lambda pos: src.count("\n", 0, pos) + 1  # noqa: E731
    funcs = [(m.group(1), m.start()) for m in re.finditer(r"async fn (\w+)\s*\(\)", src)]  # incl. the disabled _test_* ones
    cases = []
    for idx, (name, pos) in enumerate(funcs):
        end = funcs[idx + 1][1] if idx + 1 < len(funcs) else len(src)
        body = src[pos:end]
        if name == "test_all_cases" or not name.startswith("test_"):
            continue  # _test_*: disabled in the reference ("require metadata checker"), not part of test_all_cases
        lets = {}
        for m in re.finditer(r"let (\w+)\s*=\s*", body):
            # expression up to the terminating ';' at depth 0
            i, depth, instr = m.end(), 0, False
            while i < len(body):
                ch = body[i]
                if instr:
                    if ch == "\\":
                        i += 1
                    elif ch == '"':
                        instr = False
                elif ch == '"':
                    instr = True
                elif ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == ";" and depth == 0:
                    break
                i += 1
            lets[m.group(1)] = " ".join(body[m.end():i].split())
        lets.pop("table_name", None)
        lets.pop("client", None)
        steps = []
        for m in re.finditer(r"\b(init_table|execute_upsert|check_upsert|check_upsert_debug)\s*\(", body):
            args_text, _ = call_args(body, m.end() - 1)
            steps.append((m.group(1), split_top(" ".join(args_text.split())), m.start()))
        init = next(s for s in steps if s[0] == "init_table")
        schema_names = re.findall(r'"([^"]+)"', re.search(r"Schema::new\(\s*(\[[^\]]*\]|vec!\[[^\]]*\])", init[1][2]).group(1))
        pks = re.findall(r'"([^"]+)"', init[1][3])
        ranges = re.findall(r'"([^"]+)"', init[1][4])
        upserts, checks = [], []
        for kind, args, _ in steps:
            batch = eval_expr(args[0], lets)
            upserts.append(batch)
            if kind.startswith("check_upsert"):
                selected = re.findall(r'"([^"]+)"', args[2])
                flt = None
                fm = re.match(r'Some\(\s*"(.*)"\s*(?:\.to_string\(\))?\s*,?\s*\)', args[3])
                if fm:
                    flt = fm.group(1).replace('\\"', '"')
                exp_lines = re.findall(r'"([^"]*)"', args[5])
                header, rows = parse_table(exp_lines)
                checks.append({"after_upserts": len(upserts), "selected": selected, "filter": flt,
                               "expected_columns": header, "expected_rows": rows})
        # column types from the values the scenario writes
        types = {}
        for b in upserts:
            for c, vals in b.items():
                for v in vals:
                    if isinstance(v, dict):
                        types[c] = "timestamp_us"
                    elif isinstance(v, str):
                        types.setdefault(c, "utf8")
                    elif isinstance(v, int):
                        types.setdefault(c, "int32")
        text = init[1][2]
        for c in schema_names:
            if c not in types:
                types[c] = "utf8" if "DataType::Utf8" in text and "DataType::Int32" not in text else "int32"
        ups = []
        for b in upserts:
            ups.append({c: [v["ts_us"] if isinstance(v, dict) else v for v in vals] for c, vals in b.items()})
        cases.append({"name": name, "source": f"rust/lakesoul-datafusion/src/tests/upsert_tests.rs:{line_of(pos)}-{line_of(end) - 1}",
                      "schema": [[c, types[c]] for c in schema_names], "primary_keys": pks, "range_partitions": ranges,
                      "upserts": ups, "checks": checks})
    json.dump({"comment": "generated by tests/golden/make_upsert_goldens.py from the reference's upsert_tests.rs (literal inputs and "
                          "expected tables of every scenario; G9 in SURVEY.md 8c)", "cases": cases}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
