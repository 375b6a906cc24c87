"""Which nvblox:: member functions does nvblox_ros call on its mappers, and which of them does include/nvblox/ provide?

Scans the reference's ROS node sources (nvblox_ros/src/lib/*.cpp) for calls on `multi_mapper_->`, `static_mapper_->`,
`dynamic_mapper_->`, `mapper->`, including one level of accessor chaining (`static_mapper_->esdf_integrator().esdf_slice_height()`),
and looks every method name up in the mirror headers. Run where /root/reference exists; the table in INTEGRATION.md is its
output:  python tests/cpp/ros_symbol_report.py
"""
import collections
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ROS = "/root/reference/nvblox_ros/src/lib"


def mirror_methods():
    names = set()
    for path in glob.glob(os.path.join(ROOT, "include", "nvblox", "**", "*.h"), recursive=True):
        src = open(path).read()
        names |= set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\((?:[^;{}()]|\([^()]*\))*\)\s*(?:const)?\s*\{", src))
    return names


def ros_calls():
    calls = collections.defaultdict(set)
    for path in sorted(glob.glob(os.path.join(ROS, "**", "*.c*"), recursive=True)):
        src = open(path).read()
        for m in re.finditer(r"\b(multi_mapper_|static_mapper_|dynamic_mapper_|mapper)(?:->|\.)([A-Za-z_]\w*)\s*(<[^>]*>)?\(([^;]*?)\)"
                             r"(?:\.([A-Za-z_]\w*)\s*\()?(?:[^;]*?\)\.([A-Za-z_]\w*)\s*\()?", src):
            obj, meth, chained, chained2 = m.group(1), m.group(2), m.group(5), m.group(6)
            line = src.count("\n", 0, m.start()) + 1
            where = "%s:%d" % (os.path.basename(path), line)
            key = "%s::%s" % ("MultiMapper" if obj == "multi_mapper_" else "Mapper", meth)
            calls[key].add(where)
            if chained:
                calls["%s().%s" % (meth, chained)].add(where)
            if chained2:
                calls["%s().%s().%s" % (meth, chained, chained2)].add(where)
    return calls


def main():
    if not os.path.isdir(ROS):
        sys.exit("the reference tree is not on this machine")
    have = mirror_methods()
    rows = []
    for key, where in sorted(ros_calls().items()):
        leaf = re.split(r"::|\.", key.replace("()", ""))[-1]
        rows.append((key, leaf in have, sorted(where)[:3]))
    n_ok = sum(1 for r in rows if r[1])
    print("| nvblox_ros call | in include/nvblox/ | first call sites |")
    print("|---|---|---|")
    for key, ok, where in rows:
        print("| `%s` | %s | %s |" % (key, "yes" if ok else "**no**", ", ".join(where)))
    print("\n%d of %d distinct calls resolve against the mirror headers." % (n_ok, len(rows)))


if __name__ == "__main__":
    main()
