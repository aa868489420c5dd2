/* Links the packed ceremony output (rust-eth-kzg_b200/data/trusted_setup_4096.bin) into the library,
   as the reference embeds its JSON (crates/trusted_setup/src/lib.rs:5).  EKZG_TS_PATH is set by the Makefile. */
    .section .rodata
    .balign 16
    .global ekzg_trusted_setup_start
ekzg_trusted_setup_start:
    .incbin EKZG_TS_PATH
    .global ekzg_trusted_setup_end
ekzg_trusted_setup_end:
    .section .note.GNU-stack,"",@progbits
