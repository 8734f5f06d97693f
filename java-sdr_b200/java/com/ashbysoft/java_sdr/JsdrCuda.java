// JsdrCuda.java — java.lang.foreign (FFM, Java 22+) binding of libjsdrcuda.so.
//
// NOT COMPILED IN THIS REPOSITORY'S BUILD ENVIRONMENT: the image has no JDK (see
// DESIGN.md).  It is the reference-side binding a java-sdr maintainer adds next to
// fft.java / FUNcubeBPSKDemod.java (with CudaFft.java, CudaFUNcubeBPSKDemod.java and the three
// insert-only patches under ../../../patches/); every downcall below is one entry point of
// include/jsdrcuda.h, with the same argument order.  The Python mirror of these
// classes (java-sdr_b200/jsdrcuda) is what the tests drive through the same ABI.
package com.ashbysoft.java_sdr;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class JsdrCuda {
	public static final int MEM_HOST = 0, MEM_DEVICE = 1;

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
		System.getProperty("jsdrcuda.lib", "libjsdrcuda.so"), Arena.global());

	private static MethodHandle h(String name, FunctionDescriptor fd) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(
			() -> new UnsatisfiedLinkError("libjsdrcuda.so lacks " + name)), fd);
	}

	static final MethodHandle LAST_ERROR = h("jsdr_last_error", FunctionDescriptor.of(ADDRESS));
	static final MethodHandle CTX_CREATE = h("jsdr_ctx_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS));
	static final MethodHandle CTX_DESTROY = h("jsdr_ctx_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
	static final MethodHandle HOST_ALLOC = h("jsdr_host_alloc", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
	static final MethodHandle HOST_FREE = h("jsdr_host_free", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	// fft.java
	static final MethodHandle FFT_CREATE = h("jsdr_fft_create",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));
	static final MethodHandle FFT_DESTROY = h("jsdr_fft_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
	static final MethodHandle FFT_RECEIVE_F32 = h("jsdr_fft_receive_f32",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	static final MethodHandle FFT_RECEIVE_S16 = h("jsdr_fft_receive_s16",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	// FUNcubeBPSKDemod.java
	static final MethodHandle BPSK_CREATE = h("jsdr_bpsk_create",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
	static final MethodHandle BPSK_DESTROY = h("jsdr_bpsk_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
	static final MethodHandle BPSK_SET_TUNING = h("jsdr_bpsk_set_tuning",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_DOUBLE));
	static final MethodHandle BPSK_RECEIVE_F32 = h("jsdr_bpsk_receive_f32",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_INT));
	static final MethodHandle BPSK_RECEIVE_S16 = h("jsdr_bpsk_receive_s16",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT));
	static final MethodHandle BPSK_READ_BITS = h("jsdr_bpsk_read_bits",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
	static final MethodHandle PUMP_RECEIVE_S16 = h("jsdr_pump_receive_s16",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	static final MethodHandle BPSK_LAST_COUNTS = h("jsdr_bpsk_last_counts", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	static final MethodHandle BPSK_READ_DS = h("jsdr_bpsk_read_ds", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	static final MethodHandle BPSK_READ_DM = h("jsdr_bpsk_read_dm", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	static final MethodHandle BPSK_READ_FEC_COUNTERS = h("jsdr_bpsk_read_fec_counters",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
	static final MethodHandle BPSK_READ_DS_ASYNC = h("jsdr_bpsk_read_ds_async",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	static final MethodHandle BPSK_READ_COUNTERS = h("jsdr_bpsk_read_counters",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));

	static final MethodHandle BPSK_SET_AUTOTUNE = h("jsdr_bpsk_set_autotune",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));             // "FUNcube<i>-bpsk-dofft" / "-upper"
	static final MethodHandle BPSK_READ_CENTRE = h("jsdr_bpsk_read_centre",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	static final MethodHandle BPSK_SET_PRECISION = h("jsdr_bpsk_set_precision",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
	// frame stage: the sync correlator and FECDecoder.FECDecode on the device; mettab is
	// FECDecoder's own metric table (FECDecoder.java:67-100) flattened to short[512]
	static final MethodHandle BPSK_ENABLE_FEC = h("jsdr_bpsk_enable_fec",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
	static final MethodHandle BPSK_READ_FRAMES = h("jsdr_bpsk_read_frames",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
	// demod.java: FIR + NCO + detector + AGC + s16 audio in one call
	static final MethodHandle DEMOD_CREATE = h("jsdr_demod_create",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));
	static final MethodHandle DEMOD_WEIGHTS = h("jsdr_demod_weights",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT));
	static final MethodHandle DEMOD_SET_MODE = h("jsdr_demod_set_mode",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
	static final MethodHandle DEMOD_RECEIVE_AUDIO = h("jsdr_demod_receive_audio_f32",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_LONG, ADDRESS, ADDRESS, JAVA_INT));
	// waterfall.java: paintLine on the device (chained behind the fft handler's PSD)
	static final MethodHandle WATERFALL_ROWS = h("jsdr_waterfall_rows",
		FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT));

	/** Status code to message; never throws out of a handler (JavaAudio.java:321-323). */
	static String check(int rc) {
		if (rc == 0) return null;
		try {
			MemorySegment s = (MemorySegment) LAST_ERROR.invokeExact();
			return "jsdrcuda " + rc + ": " + s.reinterpret(512).getString(0);
		} catch (Throwable t) {
			return "jsdrcuda " + rc;
		}
	}

	// ---- the context every patched handler shares (one per JVM / GPU), created on first use.
	// null when libjsdrcuda.so or a CUDA device is missing: the handlers then keep running the
	// reference's own Java arithmetic (there is no CPU fallback inside the library).
	private static Context sharedCtx;
	private static boolean sharedTried;

	static synchronized Context shared(ILogger logger) {
		if (!sharedTried) {
			sharedTried = true;
			try {
				sharedCtx = new Context(Integer.getInteger("jsdrcuda.device", 0));
			} catch (Throwable t) {
				logger.statusMsg("libjsdrcuda.so not in use: " + t);
			}
		}
		return sharedCtx;
	}

	/** One context per JVM / GPU; pinned rings are allocated from it.  The header's threading
	 *  rule (a context and its handles are driven by one thread at a time) is kept by taking
	 *  `lock` around every native call: CudaFft and CudaFUNcubeBPSKDemod do, and nothing else
	 *  calls into the library. */
	public static final class Context implements AutoCloseable {
		final MemorySegment handle;
		final Arena arena = Arena.ofShared();
		final Object lock = new Object();

		public Context(int device) throws Throwable {
			MemorySegment out = arena.allocate(ADDRESS);
			String err = check((int) CTX_CREATE.invokeExact(device, out));
			if (err != null) throw new IllegalStateException(err);   // no CPU fallback by design
			handle = out.get(ADDRESS, 0);
		}

		/** Pinned host memory the audio thread fills and the GPU reads by DMA. */
		public MemorySegment pinned(long bytes) throws Throwable {
			MemorySegment out = arena.allocate(ADDRESS);
			String err = check((int) HOST_ALLOC.invokeExact(handle, bytes, out));
			if (err != null) throw new OutOfMemoryError(err);
			return out.get(ADDRESS, 0).reinterpret(bytes);
		}

		/** Returns a pinned() segment to the library (no-op for null).  Caller holds `lock`. */
		void free(MemorySegment p) {
			if (p == null) return;
			try { int rc = (int) HOST_FREE.invokeExact(handle, p); } catch (Throwable t) { }
		}

		@Override
		public void close() {
			try { int rc = (int) CTX_DESTROY.invokeExact(handle); } catch (Throwable t) { }
			arena.close();
		}
	}
}
