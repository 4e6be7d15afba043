"""Extract the reference's golden vectors for the result path into JSON fixtures that travel with the repo
(/root/reference does not exist on the GPU box).  Run in the build container:

    python tests/golden/make_golden.py [/root/reference]

Writes tests/golden/reference_fixture_cases.json: the 35 cases of src/duckdb_fixture_cases.mbt:4-262 (name, sql,
columns, expected cell strings, null mask) -- the exact VARCHAR renderings libduckdb 1.4.3 produced for the
reference's own fixture generator (scripts/generate_duckdb_fixtures.js).  Data only; no reference code is copied."""
import json
import os
import re
import sys


def parse_cases(text: str):
    body = text[text.index("= [") + 2:]
    # MoonBit record literals -> JSON: quote the keys, drop trailing commas
    body = re.sub(r"(?m)^(\s*)(name|sql|columns|rows|nulls):", r'\1"\2":', body)
    body = re.sub(r",(\s*[\]}])", r"\1", body)
    return json.loads(body.strip())


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(ref, "src", "duckdb_fixture_cases.mbt")
    cases = parse_cases(open(src, encoding="utf-8").read())
    for c in cases:
        assert set(c) == {"name", "sql", "columns", "rows", "nulls"}, c
        assert all(len(r) == len(c["columns"]) for r in c["rows"]) and len(c["rows"]) == len(c["nulls"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixture_cases.json")
    json.dump({"source": "src/duckdb_fixture_cases.mbt:4-262", "libduckdb": "@duckdb/node-api 1.4.3-r.3", "cases": cases},
              open(out, "w"), indent=1, ensure_ascii=False)
    print(f"{len(cases)} cases -> {out}")


if __name__ == "__main__":
    main()
