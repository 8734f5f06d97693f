// CudaFft.java — the GPU side of fft.java's receive(): fft.java keeps its class, its JPanel,
// its menu, its painter and its published array, and delegates the arithmetic of
// fft.java:190-224 (forward complex DFT, PSD in dB, first strict maximum, peak Hz in wrapping
// int32) to libjsdrcuda.so through this helper.  patches/fft.java.patch is the literal diff:
// one field, one line in setup(), five lines at the top of receive().
//
// Why a helper and not a replacement class: jsdr.java:476 does tabs.add("FFT", new fft(...)),
// which needs a java.awt.Component, and the painter (fft.java:86-179) reads fft's PRIVATE
// psd[] — so the original class has to stay the component and own the array; it hands that
// array to receive() below, which fills it in place.
//
// Threading: receive() runs on JavaAudio's "run" thread (JavaAudio.java:298-304), attach() on
// whichever thread publishes "audio-change" (jsdr.java:537); both hold the shared context's
// lock around every native call, as include/jsdrcuda.h requires (one thread per context at a
// time).  Nothing here throws: an exception escaping receive() would end ingest
// (JavaAudio.java:321-323), so failures are reported through ILogger.statusMsg and answered
// with `false`, which makes fft.receive() fall through to its own Java arithmetic.
//
// NOT COMPILED HERE (no JDK in the build image); see INTEGRATION.md.
package com.ashbysoft.java_sdr;

import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

final class CudaFft {
	private final JsdrCuda.Context ctx;
	private final ILogger logger;
	private final int n, rate;
	private MemorySegment handle, pinIn, pinOut, pinPeak;
	private boolean failed;

	private CudaFft(JsdrCuda.Context ctx, ILogger logger, int n, int rate) {
		this.ctx = ctx;
		this.logger = logger;
		this.n = n;
		this.rate = rate;
	}

	/** Called from fft.setup() (fft.java:63-77).  Returns the helper for this block geometry:
	 *  the old one if nothing changed, a new one (the old one's handle and pinned buffers
	 *  released) otherwise, or null when the library is not installed. */
	static CudaFft attach(CudaFft old, ILogger logger, AudioDescriptor adsc) {
		JsdrCuda.Context ctx;
		try {
			ctx = JsdrCuda.shared(logger);                       // first use loads libjsdrcuda.so (static initialiser)
		} catch (Throwable t) {                                  // library absent: ExceptionInInitializerError / NoClassDefFoundError
			logger.statusMsg("libjsdrcuda.so not in use: " + t);
			return null;
		}
		if (ctx == null) return null;
		int n = adsc.blen / adsc.size;                           // fft.java:67
		if (old != null && old.n == n && old.rate == adsc.rate && !old.failed) return old;
		synchronized (ctx.lock) {
			if (old != null) old.close();
			CudaFft f = new CudaFft(ctx, logger, n, adsc.rate);
			try {
				MemorySegment out = ctx.arena.allocate(ValueLayout.ADDRESS);
				String err = JsdrCuda.check((int) JsdrCuda.FFT_CREATE.invokeExact(ctx.handle, n, adsc.rate, 1, out));
				if (err != null) { logger.statusMsg(err); return null; }
				f.handle = out.get(ValueLayout.ADDRESS, 0);
				f.pinIn = ctx.pinned(8L * n);
				f.pinOut = ctx.pinned(4L * (n + 2));
				f.pinPeak = ctx.pinned(4);
				return f;
			} catch (Throwable t) {
				logger.statusMsg("CudaFft: " + t);
				f.close();
				return null;
			}
		}
	}

	/** fft.java:190-224 for one block.  buf: 2n floats, interleaved I,Q, caller reuses it
	 *  (JavaAudio.java:224), so it is copied into the pinned ring first (fft.java:192).
	 *  psd: fft's own float[n+2], filled in place.  false = not done, use the Java path. */
	boolean receive(float[] buf, float[] psd) {
		if (failed || psd.length != n + 2 || buf.length < 2 * n) return false;
		synchronized (ctx.lock) {
			try {
				MemorySegment.copy(buf, 0, pinIn, ValueLayout.JAVA_FLOAT, 0, 2 * n);
				String err = JsdrCuda.check((int) JsdrCuda.FFT_RECEIVE_F32.invokeExact(
					handle, pinIn, 1, pinOut, pinPeak, JsdrCuda.MEM_HOST));
				if (err != null) { fail(err); return false; }
				MemorySegment.copy(pinOut, ValueLayout.JAVA_FLOAT, 0, psd, 0, n + 2);
				return true;
			} catch (Throwable t) {
				fail("CudaFft: " + t);
				return false;
			}
		}
	}

	private void fail(String msg) {
		failed = true;                                           // stay on the Java path from now on
		logger.statusMsg(msg);
	}

	/** Caller holds ctx.lock. */
	private void close() {
		try {
			if (handle != null) { int rc = (int) JsdrCuda.FFT_DESTROY.invokeExact(handle); }
		} catch (Throwable t) { }
		handle = null;
		ctx.free(pinIn);
		ctx.free(pinOut);
		ctx.free(pinPeak);
		pinIn = pinOut = pinPeak = null;
	}
}
