/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Montgomery product for the two BN254 primes with the ADX / BMI2 instructions (mulx + the two independent carry
 * chains adcx / adox) the reference's generated routines are built from
 * (depends/ffiasm/src/montgomerybuilder.js:15-89: word-serial CIOS, one multiplication row and one reduction row per
 * limb of b, `canOptimizeConsensys` no-extra-carry-word form since the top limb of q is below 2^63 - 1 (:19), final
 * compare and single subtraction (:62-79) so the result is canonical).  Written from that algorithm, not from the
 * generated text; the plain-C restatement in field_tmpl.inc stays as the portable fallback and as its cross-check
 * (tests/test_oracle.py runs both against Python big integers).  This is what makes the CPU baseline an "ADX assembly"
 * prover like the reference, instead of a compiler-scheduled __int128 one (38 ns -> about half per product).
 *
 *   void <F>_rawMMul_adx(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]);     r may alias a or b
 */
    .intel_syntax noprefix
    .text

/* one CIOS round: t += a * b[i]; m = t0 * np; t = (t + m * q) / 2^64      t = r8..r11, A = r13, scratch rax r12 */
.macro CIOS_ROUND i, qsym, npsym
    xor     eax, eax                         /* clears CF and OF */
    mov     rdx, [rcx + 8 * \i]
    mulx    r12, rax, [rsi]
    adox    r8, rax
    mulx    r13, rax, [rsi + 8]
    adcx    r9, r12
    adox    r9, rax
    mulx    r12, rax, [rsi + 16]
    adcx    r10, r13
    adox    r10, rax
    mulx    r13, rax, [rsi + 24]
    adcx    r11, r12
    adox    r11, rax
    mov     eax, 0
    adcx    r13, rax
    adox    r13, rax                         /* A = top word of t + a * b[i] */
    mov     rdx, r8
    imul    rdx, [rip + \npsym]              /* m = t0 * (-q^-1 mod 2^64) */
    xor     eax, eax
    mulx    r12, rax, [rip + \qsym]
    adcx    rax, r8                          /* low word cancels, carry out */
    mov     r8, r12
    adcx    r8, r9
    mulx    r9, rax, [rip + \qsym + 8]
    adox    r8, rax
    adcx    r9, r10
    mulx    r10, rax, [rip + \qsym + 16]
    adox    r9, rax
    adcx    r10, r11
    mulx    r11, rax, [rip + \qsym + 24]
    adox    r10, rax
    mov     eax, 0
    adcx    r11, rax
    adox    r11, r13
.endm

.macro MMUL name, qsym, npsym
    .globl  \name
    .type   \name, @function
\name:
    push    r12
    push    r13
    mov     rcx, rdx                         /* b; rdx is mulx's implicit operand */
    xor     r8d, r8d
    xor     r9d, r9d
    xor     r10d, r10d
    xor     r11d, r11d
    CIOS_ROUND 0, \qsym, \npsym
    CIOS_ROUND 1, \qsym, \npsym
    CIOS_ROUND 2, \qsym, \npsym
    CIOS_ROUND 3, \qsym, \npsym
    /* t in [0, 2q): subtract q once if t >= q */
    mov     rax, r8
    mov     rdx, r9
    mov     r12, r10
    mov     r13, r11
    sub     rax, [rip + \qsym]
    sbb     rdx, [rip + \qsym + 8]
    sbb     r12, [rip + \qsym + 16]
    sbb     r13, [rip + \qsym + 24]
    cmovnc  r8, rax
    cmovnc  r9, rdx
    cmovnc  r10, r12
    cmovnc  r11, r13
    mov     [rdi], r8
    mov     [rdi + 8], r9
    mov     [rdi + 16], r10
    mov     [rdi + 24], r11
    pop     r13
    pop     r12
    ret
    .size   \name, . - \name
.endm

    MMUL Fq_rawMMul_adx, oracle_fq_q, oracle_fq_np
    MMUL Fr_rawMMul_adx, oracle_fr_q, oracle_fr_np

    .section .rodata
    .align 32
oracle_fq_q:  .quad 0x3c208c16d87cfd47, 0x97816a916871ca8d, 0xb85045b68181585d, 0x30644e72e131a029
oracle_fq_np: .quad 0x87d20782e4866389
    .align 32
oracle_fr_q:  .quad 0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029
oracle_fr_np: .quad 0xc2e1f593efffffff

    .section .note.GNU-stack, "", @progbits
