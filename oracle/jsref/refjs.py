"""Loads the UNMODIFIED reference sources (/root/reference/lib/*.js, the order
of build.sh: `cat header glpdebug.js lib/*.js footer`) into the minijs
interpreter.  TEST INFRASTRUCTURE; runs only in the build container -- the
reference does not exist on the GPU box and nothing here is copied into the
repo.  The MathProg translator (glpmpl*.js, glpapi14.js) and the legacy LPX
shim is loaded too: branch_drtom calls lpx_eval_tab_row (lib/glpios09.js)."""
import glob
import os

from minijs import Interp, NativeFunc, JSArray, UNDEF, js_to_str

REF = os.environ.get("GLPK_JS_REFERENCE", "/root/reference")
SKIP = ("glpmpl", "glpapi14")


def _sprintf(this, args):
    """stand-in for the php.js sprintf of lib/glpapi.js (messages only)"""
    fmt = js_to_str(args[0]) if args else ""
    out, ai, i = [], 1, 0
    while i < len(fmt):
        c = fmt[i]
        if c != "%":
            out.append(c)
            i += 1
            continue
        j = i + 1
        while j < len(fmt) and fmt[j] not in "scboxXuideEfFgG%":
            j += 1
        spec = fmt[i:j + 1]
        if spec.endswith("%"):
            out.append("%")
        else:
            v = args[ai] if ai < len(args) else UNDEF
            ai += 1
            try:
                if spec[-1] in "dixXou":
                    out.append(spec.replace("i", "d").replace("u", "d") % int(v))
                elif spec[-1] in "eEfFgG":
                    out.append(spec % float(v))
                else:
                    out.append(spec % js_to_str(v))
            except (TypeError, ValueError):
                out.append(js_to_str(v))
        i = j + 1
    return "".join(out)


def load(debug_asserts=True, quiet=True):
    I = Interp()
    I.globals["exports"] = {}
    files = [os.path.join(REF, "glpdebug.js" if debug_asserts else "glprelease.js")]
    files += [f for f in sorted(glob.glob(os.path.join(REF, "lib", "*.js")))
              if not os.path.basename(f).startswith(SKIP)]
    for f in files:
        with open(f, encoding="utf-8", errors="replace") as fh:
            I.run(fh.read(), os.path.basename(f))
    I.globals["sprintf"] = NativeFunc(_sprintf, "sprintf")
    I.api = I.globals["exports"]
    I.loaded_files = [os.path.relpath(f, REF) for f in files]
    return I


def read_lp_text(I, lp, text):
    """glp_read_lp through the reference's own character-callback reader"""
    pos = [0]

    def getc(this, args):
        if pos[0] < len(text):
            c = text[pos[0]]
            pos[0] += 1
            return c
        return -1
    return I.api["glp_read_lp"].call(I.globals, [lp, None, NativeFunc(getc, "getc")])
