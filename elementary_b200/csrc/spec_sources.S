/* spec_sources.S — embeds the device sources of K1 into libelem_b200.so (spec_host.cpp hands them to NVRTC when a voice group's
 * render program is specialised): the run-time compiled kernel is always built from exactly the text this library was built from. */
    .section .rodata
#define EMBED(sym, file) \
    .global sym ; \
    .type sym, @object ; \
sym: ; \
    .incbin file ; \
    .byte 0 ; \
    .size sym, . - sym

EMBED(eb_src_render_kernel_cu, "render_kernel.cu")
EMBED(eb_src_render_ops_inc, "render_ops.inc")
EMBED(eb_src_program_h, "program.h")
EMBED(eb_src_kernels_h, "kernels.h")
EMBED(eb_src_rtc_compat_h, "rtc_compat.h")
    .section .note.GNU-stack,"",@progbits
