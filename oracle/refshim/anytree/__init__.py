"""Test-harness stand-in for `anytree` (reference pins anytree>=2.4.3,<=2.13.0).
NOT PART OF THE PRODUCT — see oracle/refshim/cgen/__init__.py.  Provides the five names
the reference imports: NodeMixin, PostOrderIter, RenderTree, ContStyle, findall."""


class LoopError(RuntimeError):
    pass


class NodeMixin:
    separator = "/"

    @property
    def parent(self):
        return getattr(self, "_NodeMixin__parent", None)

    @parent.setter
    def parent(self, value):
        old = self.parent
        if old is value:
            return
        if value is not None:
            node = value
            while node is not None:
                if node is self:
                    raise LoopError("cannot set parent: node would be its own ancestor")
                node = node.parent
        if old is not None:
            old._NodeMixin__children_list().remove(self)
        self._NodeMixin__parent = value
        if value is not None:
            value._NodeMixin__children_list().append(self)

    def __children_list(self):
        try:
            return self._NodeMixin__children
        except AttributeError:
            self._NodeMixin__children = []
            return self._NodeMixin__children

    @property
    def children(self):
        return tuple(self.__children_list())

    @children.setter
    def children(self, children):
        for c in list(self.__children_list()):
            c.parent = None
        for c in children:
            c.parent = self

    @property
    def path(self):
        out = []
        node = self
        while node is not None:
            out.insert(0, node)
            node = node.parent
        return tuple(out)

    @property
    def ancestors(self):
        return self.path[:-1]

    @property
    def descendants(self):
        return tuple(PreOrderIter(self))[1:]

    @property
    def root(self):
        node = self
        while node.parent is not None:
            node = node.parent
        return node

    @property
    def siblings(self):
        p = self.parent
        if p is None:
            return ()
        return tuple(n for n in p.children if n is not self)

    @property
    def leaves(self):
        return tuple(n for n in PreOrderIter(self) if n.is_leaf)

    @property
    def is_leaf(self):
        return len(self.__children_list()) == 0

    @property
    def is_root(self):
        return self.parent is None

    @property
    def height(self):
        ch = self.__children_list()
        return max(c.height for c in ch) + 1 if ch else 0

    @property
    def depth(self):
        return len(self.path) - 1


def PreOrderIter(node, filter_=None, stop=None, maxlevel=None):
    def rec(n):
        yield n
        for c in n.children:
            yield from rec(c)
    for n in rec(node):
        if filter_ is None or filter_(n):
            yield n


def PostOrderIter(node, filter_=None, stop=None, maxlevel=None):
    def rec(n):
        for c in n.children:
            yield from rec(c)
        yield n
    for n in rec(node):
        if filter_ is None or filter_(n):
            yield n


def findall(node, filter_=None, stop=None, maxlevel=None, mincount=None, maxcount=None):
    return tuple(PreOrderIter(node, filter_))


class AbstractStyle:
    def __init__(self, vertical, cont, end):
        self.vertical, self.cont, self.end = vertical, cont, end
        self.empty = " " * len(end)


class ContStyle(AbstractStyle):
    def __init__(self):
        super().__init__("│   ", "├── ", "└── ")


class AsciiStyle(AbstractStyle):
    def __init__(self):
        super().__init__("|   ", "|-- ", "+-- ")


class RenderTree:
    def __init__(self, node, style=ContStyle(), childiter=list, maxlevel=None):
        if not isinstance(style, AbstractStyle):
            style = style()
        self.node, self.style, self.childiter = node, style, childiter

    def __iter__(self):
        return self._walk(self.node, ())

    def _walk(self, node, continues):
        if not continues:
            yield ("", "", node)
        else:
            st = self.style
            indent = "".join(st.vertical if c else st.empty for c in continues[:-1])
            branch = st.cont if continues[-1] else st.end
            pre = indent + branch
            fill = "".join(st.vertical if c else st.empty for c in continues)
            yield (pre, fill, node)
        children = self.childiter(node.children)
        for i, c in enumerate(children):
            yield from self._walk(c, continues + (i < len(children) - 1,))

    def by_attr(self, attrname="name"):
        lines = []
        for pre, fill, node in self:
            attr = attrname(node) if callable(attrname) else getattr(node, attrname, "")
            if isinstance(attr, (list, tuple)):
                ls = attr
            else:
                ls = str(attr).split("\n")
            lines.append(f"{pre}{ls[0]}")
            for l in ls[1:]:
                lines.append(f"{fill}{l}")
        return "\n".join(lines)

    def __str__(self):
        return "\n".join(f"{pre}{node!r}" for pre, _, node in self)
