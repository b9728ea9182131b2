"""Test-harness stand-in for the `cgen` package (pinned by the reference at
cgen>=2020.1,<2026; not installable here: no network).

NOT PART OF THE PRODUCT.  It exists only so that `/root/reference` (devito) can be
imported in the build container to (a) generate golden vectors for tests/golden and
(b) emit its own C code for the CPU baseline (oracle/_ref).  Written from the published
cgen API (a tree of `Generable`s whose `generate()` yields C source lines); only the
classes the reference imports are provided (SURVEY.md §8c lists the import sites).
"""
import numpy as np


def dtype_to_ctype(dtype):
    if dtype is None:
        raise ValueError("dtype may not be None")
    dtype = np.dtype(dtype)
    table = {
        np.dtype(np.int64): "long", np.dtype(np.uint64): "unsigned long",
        np.dtype(np.int32): "int", np.dtype(np.uint32): "unsigned int",
        np.dtype(np.int16): "short int", np.dtype(np.uint16): "short unsigned int",
        np.dtype(np.int8): "signed char", np.dtype(np.uint8): "unsigned char",
        np.dtype(np.float32): "float", np.dtype(np.float64): "double",
        np.dtype(np.bool_): "bool",
        np.dtype(np.complex64): "float _Complex", np.dtype(np.complex128): "double _Complex",
    }
    try:
        return table[dtype]
    except KeyError:
        raise ValueError(f"unable to map dtype '{dtype}'")


class Generable:
    def __str__(self):
        return "\n".join(line.rstrip() for line in self.generate())

    def generate(self, with_semicolon=True):
        raise NotImplementedError


class Declarator(Generable):
    def generate(self, with_semicolon=True):
        tp_lines, tp_decl = self.get_decl_pair()
        tp_lines = list(tp_lines)
        yield from tp_lines[:-1]
        sc = ";" if with_semicolon else ""
        if tp_decl is None:
            yield f"{tp_lines[-1]}{sc}"
        else:
            yield f"{tp_lines[-1]} {tp_decl}{sc}"

    def get_decl_pair(self):
        raise NotImplementedError

    def inline(self, with_semicolon=False):
        tp_lines, tp_decl = self.get_decl_pair()
        tp_lines = " ".join(tp_lines)
        if tp_decl is None:
            return tp_lines
        return f"{tp_lines} {tp_decl}"


class POD(Declarator):
    def __init__(self, dtype, name):
        self.dtype = np.dtype(dtype)
        self.name = name

    def get_decl_pair(self):
        return [dtype_to_ctype(self.dtype)], self.name


class Value(Declarator):
    def __init__(self, typename, name):
        self.typename = typename
        self.name = name

    def get_decl_pair(self):
        return [self.typename], self.name


class NestedDeclarator(Declarator):
    def __init__(self, subdecl):
        self.subdecl = subdecl

    @property
    def name(self):
        return self.subdecl.name

    def get_decl_pair(self):
        return self.subdecl.get_decl_pair()


class DeclSpecifier(NestedDeclarator):
    def __init__(self, subdecl, spec, sep=" "):
        super().__init__(subdecl)
        self.spec = spec
        self.sep = sep

    def get_decl_pair(self):
        def add_spec(sub_it):
            it = iter(sub_it)
            try:
                yield f"{self.spec}{self.sep}{next(it)}"
            except StopIteration:
                pass
            yield from it
        sub_tp, sub_decl = self.subdecl.get_decl_pair()
        return add_spec(sub_tp), sub_decl


class Static(DeclSpecifier):
    def __init__(self, subdecl):
        super().__init__(subdecl, "static")


class Const(NestedDeclarator):
    def get_decl_pair(self):
        sub_tp, sub_decl = self.subdecl.get_decl_pair()
        return sub_tp, f"const {sub_decl}"


class Pointer(NestedDeclarator):
    def get_decl_pair(self):
        sub_tp, sub_decl = self.subdecl.get_decl_pair()
        return sub_tp, f"*{sub_decl}"


class AlignedAttribute(NestedDeclarator):
    def __init__(self, align_bytes, subdecl):
        super().__init__(subdecl)
        self.align_bytes = align_bytes

    def get_decl_pair(self):
        sub_tp, sub_decl = self.subdecl.get_decl_pair()
        return sub_tp, f"{sub_decl} __attribute__ ((aligned ({self.align_bytes})))"


class Extern(DeclSpecifier):
    def __init__(self, language, subdecl):
        self.language = language
        super().__init__(subdecl, f'extern "{language}"')


class Template(NestedDeclarator):
    def __init__(self, template_spec, subdecl):
        super().__init__(subdecl)
        self.template_spec = template_spec

    def generate(self, with_semicolon=False):
        yield f"template <{self.template_spec}>"
        yield from self.subdecl.generate(with_semicolon=with_semicolon)


class FunctionDeclaration(NestedDeclarator):
    def __init__(self, subdecl, arg_decls):
        super().__init__(subdecl)
        self.arg_decls = arg_decls

    def get_decl_pair(self):
        sub_tp, sub_decl = self.subdecl.get_decl_pair()
        return sub_tp, "{}({})".format(
            sub_decl, ", ".join(ad.inline() for ad in self.arg_decls))


class Struct(Declarator):
    def __init__(self, tpname, fields, declname=None, pad_bytes=0):
        self.tpname = tpname
        self.fields = fields
        self.declname = declname
        self.pad_bytes = pad_bytes

    def get_decl_pair(self):
        def get_tp():
            if self.tpname is not None:
                yield f"struct {self.tpname}"
            else:
                yield "struct"
            yield "{"
            for f in self.fields:
                for f_line in f.generate():
                    yield "  " + f_line
            if self.pad_bytes:
                yield f"  unsigned char _cgen_pad[{self.pad_bytes}];"
            yield "}"
        return get_tp(), self.declname


class Define(Generable):
    def __init__(self, symbol, value):
        self.symbol = symbol
        self.value = value

    def generate(self, with_semicolon=True):
        yield f"#define {self.symbol} {self.value}"


class Include(Generable):
    def __init__(self, filename, system=True):
        self.filename = filename
        self.system = system

    def generate(self, with_semicolon=True):
        if self.system:
            yield f"#include <{self.filename}>"
        else:
            yield f'#include "{self.filename}"'


class Pragma(Generable):
    def __init__(self, value):
        self.value = value

    def generate(self, with_semicolon=True):
        yield f"#pragma {self.value}"


class Statement(Generable):
    def __init__(self, text):
        self.text = text

    def generate(self, with_semicolon=True):
        yield self.text + ";"


class ExpressionStatement(Statement):
    pass


class Assign(Generable):
    def __init__(self, lvalue, rvalue):
        self.lvalue = lvalue
        self.rvalue = rvalue

    def generate(self, with_semicolon=True):
        yield f"{self.lvalue} = {self.rvalue};"


class Line(Generable):
    def __init__(self, text=""):
        self.text = text

    def generate(self, with_semicolon=True):
        yield self.text


class Comment(Generable):
    def __init__(self, text, skip_space=False):
        self.text = text
        self.fmt = "/*{comment}*/" if skip_space else "/* {comment} */"

    def generate(self, with_semicolon=True):
        yield self.fmt.format(comment=self.text)


class MultilineComment(Generable):
    def __init__(self, text, skip_space=False):
        self.text = text
        self.skip_space = skip_space

    def generate(self, with_semicolon=True):
        yield "/**"
        line_begin = " *" if self.skip_space else " * "
        for line in self.text.splitlines():
            yield line_begin + line
        yield " */"


class LineComment(Generable):
    def __init__(self, text):
        self.text = text

    def generate(self, with_semicolon=True):
        yield f"// {self.text}"


class Initializer(Generable):
    def __init__(self, vdecl, data):
        self.vdecl = vdecl
        self.data = data

    def generate(self, with_semicolon=True):
        tp_lines, tp_decl = self.vdecl.get_decl_pair()
        tp_lines = list(tp_lines)
        yield from tp_lines[:-1]
        sc = ";" if with_semicolon else ""
        yield f"{tp_lines[-1]} {tp_decl} = {self.data}{sc}"


class Block(Generable):
    def __init__(self, contents=None):
        contents = [] if contents is None else contents
        if isinstance(contents, Block):
            contents = contents.contents
        self.contents = list(contents)

    def generate(self, with_semicolon=True):
        yield "{"
        for item in self.contents:
            for item_line in item.generate():
                yield "  " + item_line
        yield "}"

    def append(self, data):
        self.contents.append(data)

    def extend(self, data):
        self.contents.extend(data)

    def insert(self, i, data):
        self.contents.insert(i, data)


class Collection(Block):
    def generate(self, with_semicolon=True):
        for c in self.contents:
            yield from c.generate()


Module = Collection


class FunctionBody(Generable):
    def __init__(self, fdecl, body):
        self.fdecl = fdecl
        self.body = body

    def generate(self, with_semicolon=True):
        yield from self.fdecl.generate(with_semicolon=False)
        yield from self.body.generate()


class If(Generable):
    def __init__(self, condition, then_, else_=None):
        self.condition = condition
        self.then_ = then_
        self.else_ = else_

    def generate(self, with_semicolon=True):
        cond = str(self.condition)
        yield f"if ({cond})"
        if isinstance(self.then_, Block):
            yield from self.then_.generate()
        else:
            for line in self.then_.generate():
                yield "  " + line
        if self.else_ is not None:
            yield "else"
            if isinstance(self.else_, Block):
                yield from self.else_.generate()
            else:
                for line in self.else_.generate():
                    yield "  " + line


class Loop(Generable):
    def __init__(self, body):
        self.body = body

    def intro_line(self):
        raise NotImplementedError

    def outro_line(self):
        return None

    def generate(self, with_semicolon=True):
        il = self.intro_line()
        if il is not None:
            yield il
        if isinstance(self.body, Block):
            yield from self.body.generate()
        else:
            for line in self.body.generate():
                yield "  " + line
        ol = self.outro_line()
        if ol is not None:
            yield ol


class While(Loop):
    def __init__(self, condition, body):
        self.condition = condition
        super().__init__(body)

    def intro_line(self):
        return f"while ({self.condition})"


class For(Loop):
    def __init__(self, start, condition, update, body):
        self.start = start
        self.condition = condition
        self.update = update
        super().__init__(body)

    def intro_line(self):
        return f"for ({self.start}; {self.condition}; {self.update})"


class IfDef(Module):
    def __init__(self, condition, iflines, elselines):
        ifdef_line = Line(f"#ifdef {condition}")
        if len(elselines):
            lines = [ifdef_line] + list(iflines) + [Line("#else")] + list(elselines) \
                + [Line("#endif")]
        else:
            lines = [ifdef_line] + list(iflines) + [Line("#endif")]
        super().__init__(lines)


class IfNDef(Module):
    def __init__(self, condition, ifndeflines, elselines):
        ifndef_line = Line(f"#ifndef {condition}")
        if len(elselines):
            lines = [ifndef_line] + list(ifndeflines) + [Line("#else")] \
                + list(elselines) + [Line("#endif")]
        else:
            lines = [ifndef_line] + list(ifndeflines) + [Line("#endif")]
        super().__init__(lines)
