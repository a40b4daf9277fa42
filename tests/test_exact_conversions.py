"""K_A replaces the I2F / F2F conversions of the 8-bit footprint taps by magic-number constructions on the
FMA / FP64 pipes (photobundle_b200/csrc/k_step.cu: u8_to_f32, u8_to_f64, half_diff_f32, diff_f64).  The
claim "exact for every input" is checked here exhaustively with the same bit manipulations in C on the
host (the GPU parity tests check the end result: bit-identical residuals)."""
import subprocess
import textwrap


def test_magic_number_conversions_are_exact(tmp_path):
    src = tmp_path / "magic.c"
    src.write_text(textwrap.dedent(r'''
        #include <stdio.h>
        #include <string.h>
        #include <stdint.h>
        #include <math.h>
        static float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
        static double hilo(int32_t hi, int32_t lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
        int main(void) {
          int bad = 0;
          for (int b = 0; b < 256; ++b) {
            if (i2f(0x4B000000 | b) - 8388608.0f != (float)b) ++bad;                       /* u8_to_f32 */
            if (hilo(0x43300000, b) - 4503599627370496.0 != (double)b) ++bad;              /* u8_to_f64 */
          }
          for (int d = -255; d <= 255; ++d) {
            if (fmaf(i2f(0x4B400000 + d), 0.5f, -6291456.0f) != 0.5f * (float)d) ++bad;    /* half_diff_f32 */
            if (hilo(0x43300000, d ^ (int32_t)0x80000000) - (4503599627370496.0 + 2147483648.0) != (double)d) ++bad;   /* diff_f64 */
          }
          printf("%d\n", bad);
          return bad != 0;
        }
    '''))
    exe = tmp_path / "magic"
    subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0"
