"""Stand-in for codepy.toolchain: a record of compiler/linker settings + command line."""
import subprocess


def call_capture_output(cmdline, cwd=None, error_on_nonzero=True):
    try:
        p = subprocess.Popen(cmdline, cwd=cwd, stdin=subprocess.PIPE,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        stdout, stderr = p.communicate()
    except OSError as e:
        raise RuntimeError(f"error invoking '{' '.join(cmdline)}': {e}") from e
    if p.returncode and error_on_nonzero:
        raise RuntimeError(f"status {p.returncode} invoking '{' '.join(cmdline)}': "
                           f"{stderr.decode(errors='replace')}")
    return p.returncode, stdout, stderr


class CompileError(Exception):
    def __init__(self, msg, command, stdout=None, stderr=None):
        super().__init__(msg)
        self.command, self.stdout, self.stderr = command, stdout, stderr


class Toolchain:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


class GCCLikeToolchain(Toolchain):
    def get_version(self):
        _, stdout, _ = call_capture_output([self.cc, "--version"])
        return stdout

    def _cmdline(self, files, object=False):
        if object:
            ld_options = ['-c']
            link = []
        else:
            ld_options = list(self.ldflags)
            link = [f"-L{ld}" for ld in self.library_dirs]
            link.extend(f"-l{lib}" for lib in self.libraries)
        return ([self.cc] + list(self.cflags) + ld_options
                + [f"-D{d}" for d in self.defines]
                + [f"-U{d}" for d in self.undefines]
                + [f"-I{i}" for i in self.include_dirs]
                + list(files) + link)

    def build_extension(self, ext_file, source_files, debug=False):
        cc_cmdline = self._cmdline(source_files, False) + ["-o", ext_file]
        if debug:
            print(" ".join(cc_cmdline))
        result, stdout, stderr = call_capture_output(cc_cmdline, error_on_nonzero=False)
        if result != 0:
            raise CompileError("module compilation failed", cc_cmdline,
                               stdout.decode(errors='replace'),
                               stderr.decode(errors='replace'))


class GCCToolchain(GCCLikeToolchain):
    pass
