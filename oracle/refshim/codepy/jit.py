"""Stand-in for codepy.jit.compile_from_string: write the source, compile it to
`<name>.so` with the toolchain's command line, report whether a recompile happened."""
import os


def compile_from_string(toolchain, name, source_string, source_name="module.cpp",
                        cache_dir=None, debug=False, wait_on_error=None,
                        debug_recompile=True, object=False, source_is_binary=False,
                        sleep_delay=1):
    if isinstance(source_name, (list, tuple)):
        source_name = source_name[0]
    src = str(source_name)
    if not os.path.isabs(src):
        src = os.path.join(os.path.dirname(str(name)), src)
    ext_file = str(name) + (toolchain.o_ext if object else toolchain.so_ext)
    os.makedirs(os.path.dirname(ext_file) or ".", exist_ok=True)
    mode = "wb" if source_is_binary else "w"
    with open(src, mode) as f:
        f.write(source_string)
    if object:
        cmd = toolchain._cmdline([src], True) + ["-o", ext_file]
        from codepy.toolchain import call_capture_output, CompileError
        rc, out, err = call_capture_output(cmd, error_on_nonzero=False)
        if rc:
            raise CompileError("compilation failed", cmd, out.decode(), err.decode())
    else:
        toolchain.build_extension(ext_file, [src], debug=debug)
    return None, name, ext_file, True
