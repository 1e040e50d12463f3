#!/usr/bin/env python
"""JDK-free structural check of java/ (the Panama-FFM shim) against the reference tree.

The image has no javac, so the shim cannot be compiled.  This script does the part of a compiler's job that matters
for a drop-in: it extracts every use the reference's sources make of the classes the shim replaces --
`new T(args)`, `T.staticMember`, and `x.member(...)` for every field / parameter / local `x` declared with one of
those types -- and asserts that the shim declares a public member of that name (constructors: with that many
parameters).  It also checks that every C entry point the shim binds exists in include/kmersreps.h with the same
number of parameters.

    python tools/check_java_shim.py [/root/reference]      -> exit 0 and a summary, or exit 1 listing what is missing
Run by tests/test_java_shim.py (with the reference tree when present, else against the committed use list
tests/golden/java_member_uses.json, which this script writes with --write-golden)."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java")
GOLDEN = os.path.join(ROOT, "tests", "golden", "java_member_uses.json")
# simple name -> shim source (classes whose upstream versions the shim replaces)
CLASSES = {
    "SequenceKmers": "org/theseed/sequence/SequenceKmers.java",
    "ProteinKmers": "org/theseed/sequence/ProteinKmers.java",
    "RepGenome": "org/theseed/proteins/kmers/reps/RepGenome.java",
    "RepGenomeDb": "org/theseed/proteins/kmers/reps/RepGenomeDb.java",
    "Representation": "org/theseed/proteins/kmers/reps/RepGenomeDb.java",
    "KmerCollectionGroup": "org/theseed/proteins/kmers/KmerCollectionGroup.java",
}
PARENT = {"ProteinKmers": "SequenceKmers", "RepGenome": "ProteinKmers"}
OBJECT_MEMBERS = {"equals", "hashCode", "toString", "getClass"}


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: re.sub(r"[^\n]", " ", m.group(0)), src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    return re.sub(r"'(?:\\.|[^'\\])'", "' '", src)


def count_args(src, open_paren):
    """number of top-level arguments of the call / parameter list whose '(' is at open_paren (commas inside nested
    brackets or inside the <...> of a generic type do not count)"""
    depth, angle, i, args, seen = 0, 0, open_paren, 0, False
    while i < len(src):
        c = src[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                return args + (1 if seen else 0)
        elif c == "<" and depth == 1 and i > 0 and (src[i - 1].isalnum() or src[i - 1] == "?"):
            angle += 1
        elif c == ">" and angle and src[i - 1] != "-":
            angle -= 1
        elif depth == 1 and angle == 0 and c == ",":
            args += 1
            seen = False
        elif depth >= 1 and not c.isspace():
            seen = True
        i += 1
    return args


def reference_uses(ref_root):
    """[{file, line, cls, member, kind: 'new'|'static'|'call'|'field', nargs}] for every use found"""
    uses = []
    src_root = os.path.join(ref_root, "src")
    type_re = "|".join(sorted(CLASSES, key=len, reverse=True))
    for dp, _, fs in os.walk(src_root):
        for f in fs:
            if not f.endswith(".java"):
                continue
            path = os.path.join(dp, f)
            raw = open(path, encoding="utf-8", errors="replace").read()
            src = strip_comments(raw)
            rel = os.path.relpath(path, ref_root)

            def line_of(pos):
                return src.count("\n", 0, pos) + 1

            # constructors: new T(...) and outer.new Representation(...)
            for m in re.finditer(r"\bnew\s+(?:RepGenomeDb\s*\.\s*)?(%s)\s*\(" % type_re, src):
                uses.append({"file": rel, "line": line_of(m.start()), "cls": m.group(1), "member": "<init>", "kind": "new",
                             "nargs": count_args(src, m.end() - 1)})
            # static members: T.member
            for m in re.finditer(r"(?<![\w.])(%s)\s*\.\s*([A-Za-z_]\w*)\s*(\()?" % type_re, src):
                cls, mem = m.group(1), m.group(2)
                if mem in ("class", "Representation", "this", "new") or src[max(0, m.start() - 7):m.start()].strip().endswith("import"):
                    continue
                uses.append({"file": rel, "line": line_of(m.start()), "cls": cls, "member": mem, "kind": "static" if m.group(3) else "field",
                             "nargs": count_args(src, m.end() - 1) if m.group(3) else None})
            # variables declared with one of the types, resolved per method: a class body is cut into its top-level members
            # (brace depth 1 -> 2); a member sees the class's fields unless it declares the same name itself (with any type)
            decl_re = re.compile(r"(?<![\w.])((?:[A-Z]\w*\s*\.\s*)*[A-Z]\w*|var)(?:\s*<[^;(){}=]*?>)?(?:\s*\[\s*\])*\s+([a-z_]\w*)\s*(?=[=;,):])")

            def tracked(type_text):
                simple = re.split(r"\s*\.\s*", type_text)[-1]
                return simple if simple in CLASSES else None

            depth, chunks, start, last_end = 0, [], None, 0
            fields_text = []
            for i, c in enumerate(src):
                if c == "{":
                    depth += 1
                    if depth == 2:
                        hdr = max(last_end, src.rfind(";", last_end, i) + 1, src.rfind("}", last_end, i) + 1, src.rfind("{", last_end, i) + 1)
                        fields_text.append(src[last_end:hdr])
                        start = hdr
                elif c == "}":
                    depth -= 1
                    if depth == 1 and start is not None:
                        chunks.append((start, i + 1))
                        last_end, start = i + 1, None
            fields_text.append(src[last_end:])
            fields = {}
            for m in decl_re.finditer(" ".join(fields_text)):
                fields[m.group(2)] = tracked(m.group(1))
            for a, b in chunks:
                body = src[a:b]
                local = {}
                for m in decl_re.finditer(body):
                    t = tracked(m.group(1))
                    if m.group(2) not in local or t:
                        local[m.group(2)] = t
                for m in re.finditer(r"(?<![\w.])(this\s*\.\s*)?([a-z_]\w*)\s*\.\s*([A-Za-z_]\w*)\s*\(", body):
                    var = m.group(2)
                    t = fields.get(var) if (m.group(1) or var not in local) else local[var]
                    if t:
                        uses.append({"file": rel, "line": line_of(a + m.start()), "cls": t, "member": m.group(3), "kind": "call",
                                     "nargs": count_args(body, m.end() - 1)})
    return uses


def shim_members():
    """class -> {name -> set(parameter counts)} for public/protected members, (parent chain applied by the caller)"""
    out = {}
    for cls, rel in CLASSES.items():
        src = strip_comments(open(os.path.join(JAVA, rel), encoding="utf-8").read())
        if cls == "Representation":
            a = src.index("class Representation")
            depth, i = 0, src.index("{", a)
            start = i
            while True:
                depth += src[i] == "{"
                depth -= src[i] == "}"
                if depth == 0:
                    break
                i += 1
            body = src[start:i]
        else:
            body = src
            if cls == "RepGenomeDb":      # without the inner class
                a = src.index("class Representation")
                depth, i = 0, src.index("{", a)
                s0 = i
                while True:
                    depth += src[i] == "{"
                    depth -= src[i] == "}"
                    if depth == 0:
                        break
                    i += 1
                body = src[:a] + src[i:]
        members = {}
        for m in re.finditer(r"\b(public|protected)\s+((?:static\s+|final\s+|abstract\s+|synchronized\s+)*)(?:<[^>]+>\s*)?([\w.<>\[\], ?]+?)\s+(\w+)\s*\(", body):
            name = m.group(4)
            members.setdefault(name, set()).add(count_args(body, m.end() - 1))
        for m in re.finditer(r"\b(public)\s+%s\s*\(" % cls, body):      # constructors
            members.setdefault("<init>", set()).add(count_args(body, m.end() - 1))
        for m in re.finditer(r"\bpublic\s+static\s+final\s+[\w.<>\[\]]+\s+(\w+)\s*=", body):
            members.setdefault(m.group(1), set()).add(None)
        # varargs: a parameter list ending in "..." accepts any count >= its fixed part
        for m in re.finditer(r"\bpublic\s+(?:[\w.<>\[\]]+\s+)?(\w+)\s*\(([^)]*\.\.\.[^)]*)\)", body):
            name = "<init>" if m.group(1) == cls else m.group(1)
            fixed = m.group(2).count(",")
            members.setdefault(name, set()).update(range(fixed, fixed + 8))
        out[cls] = members
    return out


def check(uses):
    members = shim_members()
    missing = []
    for u in uses:
        cls = u["cls"]
        if u["kind"] != "new" and u["member"] in OBJECT_MEMBERS and u["kind"] == "call":
            pass
        found, c = False, cls
        while c and not found:
            ms = members[c].get(u["member"])
            if ms is not None and (u["kind"] == "field" or u["nargs"] is None or None in ms or u["nargs"] in ms):
                found = True
            c = PARENT.get(c) if u["kind"] != "new" else None
        if not found and u["kind"] == "call" and u["member"] in OBJECT_MEMBERS:
            found = True
        if not found:
            missing.append(u)
    return missing


def check_ffm_against_header():
    """every MethodHandle of KmersRepsNative.java names a function of include/kmersreps.h with the same parameter count"""
    hdr = strip_comments(open(os.path.join(ROOT, "include", "kmersreps.h")).read())
    decl = {}
    for m in re.finditer(r"\b(kr_\w+)\s*\(", hdr):
        decl.setdefault(m.group(1), count_args(hdr, m.end() - 1))
    for name, n in list(decl.items()):
        seg = hdr[hdr.index(name + "("):]
        if re.match(r"%s\s*\(\s*void\s*\)" % name, seg.replace("\n", " ")):
            decl[name] = 0
    src = strip_comments(open(os.path.join(JAVA, "org/theseed/sequence/KmersRepsNative.java")).read())
    src_raw = open(os.path.join(JAVA, "org/theseed/sequence/KmersRepsNative.java")).read()
    bad = []
    for m in re.finditer(r'h\("(kr_\w+)",\s*FunctionDescriptor\.(ofVoid|of)\(', src_raw):
        name, kind = m.group(1), m.group(2)
        n = count_args(src_raw, m.end() - 1) - (0 if kind == "ofVoid" else 1)
        if name not in decl:
            bad.append("%s is bound by the shim but not declared in include/kmersreps.h" % name)
        elif decl[name] != n:
            bad.append("%s: header declares %d parameters, the shim's FunctionDescriptor has %d" % (name, decl[name], n))
    return bad


def check_brackets():
    """every shim source has balanced (), [] and {} outside comments and literals (the one syntax check a script can do)"""
    bad = []
    pairs = {")": "(", "]": "[", "}": "{"}
    for dp, _, fs in os.walk(JAVA):
        for f in fs:
            if not f.endswith(".java"):
                continue
            src = strip_comments(open(os.path.join(dp, f), encoding="utf-8").read())
            stack = []
            for i, ch in enumerate(src):
                if ch in "([{":
                    stack.append((ch, i))
                elif ch in ")]}":
                    if not stack or stack[-1][0] != pairs[ch]:
                        bad.append("%s: unmatched '%s' at line %d" % (f, ch, src.count("\n", 0, i) + 1))
                        break
                    stack.pop()
            else:
                if stack:
                    bad.append("%s: unclosed '%s' from line %d" % (f, stack[-1][0], src.count("\n", 0, stack[-1][1]) + 1))
    return bad


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ref = args[0] if args else "/root/reference"
    if os.path.isdir(os.path.join(ref, "src")):
        uses = reference_uses(ref)
        if "--write-golden" in sys.argv:
            json.dump(uses, open(GOLDEN, "w"), indent=0, sort_keys=True)
    else:
        uses = json.load(open(GOLDEN))
    missing = check(uses)
    bad = check_ffm_against_header() + check_brackets()
    by = {}
    for u in uses:
        by.setdefault((u["cls"], u["member"]), 0)
        by[(u["cls"], u["member"])] += 1
    print("%d uses of %d distinct members of %d classes in the reference; %d unresolved; %d FFM signature problems"
          % (len(uses), len(by), len({u['cls'] for u in uses}), len(missing), len(bad)))
    for u in missing:
        print("MISSING %s.%s (%s, %s args) used at %s:%d" % (u["cls"], u["member"], u["kind"], u["nargs"], u["file"], u["line"]))
    for b in bad:
        print("FFM", b)
    return 1 if missing or bad else 0


if __name__ == "__main__":
    sys.exit(main())
