// CudaFUNcubeBPSKDemod.java — the GPU side of FUNcubeBPSKDemod.receive(): the reference class
// keeps its JPanel, its menu, its hot keys, its painter and its config / publish keys, and
// delegates the arithmetic of FUNcubeBPSKDemod.java:358-595 (tuner mix, 27-tap decimator,
// 1200 Hz VCO, 65-tap matched filter, bit timing, differential decision, sync correlator) and
// FECDecoder.FECDecode (FECDecoder.java:703-852) to libjsdrcuda.so through this helper.
// patches/FUNcubeBPSKDemod.java.patch is the literal diff: one field, one line in setup(),
// five lines at the top of receive() and the method cudaFetch() that copies the results below
// into the private fields the painter reads.  jsdr.java is not touched: it goes on doing
// tabs.add(nm, new FUNcubeBPSKDemod(fc, ...)) (jsdr.java:479-483) and dispatching hot keys.
//
// One helper = one reference instance = a bank of ONE tuner (the reference builds
// "jsdr-funcube-demods" independent instances, each with its own tuning and its own
// dofft/upper switches, :195-200); a server that wants many tuners on one stream uses the
// batched ABI directly (jsdr_bpsk_create with nchan tunings, chan_stride 0).
//
// Threading: every native call on a context happens under JsdrCuda.Context.lock (the header's
// rule: one thread per context at a time).  Menu actions (:174-189) arrive on the Swing EDT and
// only change Java fields, as in the reference; receive() — on JavaAudio's "run" thread — sees
// the new tuning / dofft / upper values as arguments and applies them before the block, so a
// retune can never land between a block's phase replay and its data kernels.  attach() keeps
// the native handle (and with it tuPhase, the filter histories and the bit-timing state) when
// the block geometry is unchanged, as the reference's setup() keeps its instance fields, and
// otherwise releases the old handle and its pinned buffers before allocating new ones.
//
// NOT COMPILED HERE (no JDK in the build image); see INTEGRATION.md.
package com.ashbysoft.java_sdr;

import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

final class CudaFUNcubeBPSKDemod {
	private static final int MAX_FRAMES = 4;
	private final JsdrCuda.Context ctx;
	private final ILogger logger;
	private final int rate, samples, maxDs, maxBits;
	private MemorySegment handle, pinIn, pinDs, pinDm, pinBits, pinBitAt, pinNbits, pinCnt, pinFrames, pinMisc;
	private double lastTuning = Double.NaN;
	private int lastDoFFT = -1, lastDoUp = -1;
	private boolean failed;

	// ---- results of the last receive(), read by FUNcubeBPSKDemod.cudaFetch()
	int nDs;                    // 9600 S/s samples produced by this block
	long firstDs;               // cntDS before this block (bitAt counts from the start of the stream)
	double[] ds, dm;            // interleaved I,Q: RxDownSample output (:486), matched filter output (:530-531)
	int nBits;
	byte[] bits;                // +1 / -1 as written to dmFECCorr (:554)
	long[] bitAt;
	long cntRaw, cntDS, cntBit, cntFEC, cntDec;   // :114, painted at :220
	int frameErrors;            // FECDecode's return for the last frame of this block, Integer.MIN_VALUE: none
	final byte[] frame = new byte[256];
	int centreBin;              // :456

	private CudaFUNcubeBPSKDemod(JsdrCuda.Context ctx, ILogger logger, int rate, int samples) {
		this.ctx = ctx;
		this.logger = logger;
		this.rate = rate;
		this.samples = samples;
		this.maxDs = samples * 9600 / rate + 2;
		this.maxBits = maxDs / 4 + 2;        // decisions are at least 5 samples apart (:537,577-578)
		ds = new double[2 * maxDs];
		dm = new double[2 * maxDs];
		bits = new byte[maxBits];
		bitAt = new long[maxBits];
	}

	/** Called from FUNcubeBPSKDemod.setup() (:192-209).  mettab = FECDecoder's own metric table. */
	static CudaFUNcubeBPSKDemod attach(CudaFUNcubeBPSKDemod old, ILogger logger, AudioDescriptor adsc, int[][] mettab) {
		JsdrCuda.Context ctx;
		try {
			ctx = JsdrCuda.shared(logger);                       // first use loads libjsdrcuda.so (static initialiser)
		} catch (Throwable t) {                                  // library absent: ExceptionInInitializerError / NoClassDefFoundError
			logger.statusMsg("libjsdrcuda.so not in use: " + t);
			return null;
		}
		if (ctx == null) return null;
		int samples = adsc.blen / adsc.size;                                         // :194
		if (old != null && old.rate == adsc.rate && old.samples == samples && !old.failed) return old;
		synchronized (ctx.lock) {
			if (old != null) old.close();
			CudaFUNcubeBPSKDemod b = new CudaFUNcubeBPSKDemod(ctx, logger, adsc.rate, samples);
			try {
				MemorySegment out = ctx.arena.allocate(ValueLayout.ADDRESS);
				MemorySegment tun = ctx.arena.allocate(ValueLayout.JAVA_DOUBLE);
				tun.set(ValueLayout.JAVA_DOUBLE, 0, 12000.0);                        // replaced by the first receive()
				String err = JsdrCuda.check((int) JsdrCuda.BPSK_CREATE.invokeExact(
					ctx.handle, adsc.rate, 1, tun, samples, out));
				if (err != null) { logger.statusMsg(err); return null; }
				b.handle = out.get(ValueLayout.ADDRESS, 0);
				MemorySegment met = ctx.arena.allocate(ValueLayout.JAVA_SHORT, 512);
				for (int r = 0; r < 2; r++)
					for (int i = 0; i < 256; i++)
						met.setAtIndex(ValueLayout.JAVA_SHORT, r * 256 + i, (short) mettab[r][i]);
				err = JsdrCuda.check((int) JsdrCuda.BPSK_ENABLE_FEC.invokeExact(b.handle, met, MAX_FRAMES));
				if (err != null) { logger.statusMsg(err); b.close(); return null; }
				b.pinIn = ctx.pinned(8L * samples);
				b.pinDs = ctx.pinned(16L * b.maxDs);
				b.pinDm = ctx.pinned(16L * b.maxDs);
				b.pinBits = ctx.pinned(b.maxBits);
				b.pinBitAt = ctx.pinned(8L * b.maxBits);
				b.pinNbits = ctx.pinned(4);
				b.pinCnt = ctx.pinned(8L * 6);                                      // 4 counters + cntFEC + cntDec
				b.pinFrames = ctx.pinned(MAX_FRAMES * (4 + 8 + 4 + 256L) + 8);
				b.pinMisc = ctx.pinned(16);
				return b;
			} catch (Throwable t) {
				logger.statusMsg("CudaFUNcubeBPSKDemod: " + t);
				b.close();
				return null;
			}
		}
	}

	/** FUNcubeBPSKDemod.receive(buf) (:358-364) for one block: doBufferTune (:366-379) or, with
	 *  doFFT, doBufferFFT (:406-464).  false = not done, the caller runs its Java path. */
	boolean receive(float[] buf, boolean doFFT, boolean doUp, double tuning) {
		if (failed || buf.length < 2 * samples) return false;
		synchronized (ctx.lock) {
			try {
				String err = null;
				if (tuning != lastTuning) {                                          // actionPerformed :174-189
					err = JsdrCuda.check((int) JsdrCuda.BPSK_SET_TUNING.invokeExact(handle, 0, tuning));
					lastTuning = tuning;
				}
				if (err == null && ((doFFT ? 1 : 0) != lastDoFFT || (doUp ? 1 : 0) != lastDoUp)) {
					err = JsdrCuda.check((int) JsdrCuda.BPSK_SET_AUTOTUNE.invokeExact(handle, doFFT ? 1 : 0, doUp ? 1 : 0));
					lastDoFFT = doFFT ? 1 : 0;
					lastDoUp = doUp ? 1 : 0;
				}
				if (err == null) {
					MemorySegment.copy(buf, 0, pinIn, ValueLayout.JAVA_FLOAT, 0, 2 * samples);
					err = JsdrCuda.check((int) JsdrCuda.BPSK_RECEIVE_F32.invokeExact(
						handle, pinIn, samples, 0L, JsdrCuda.MEM_HOST));
				}
				if (err == null) err = JsdrCuda.check((int) JsdrCuda.BPSK_LAST_COUNTS.invokeExact(handle, pinMisc));
				if (err != null) { fail(err); return false; }
				nDs = Math.min(pinMisc.get(ValueLayout.JAVA_INT, 0), maxDs);
				err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_DS.invokeExact(handle, pinDs, JsdrCuda.MEM_HOST));
				if (err == null) err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_DM.invokeExact(handle, pinDm, JsdrCuda.MEM_HOST));
				if (err == null) err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_BITS.invokeExact(
					handle, pinBits, pinBitAt, pinNbits, maxBits, JsdrCuda.MEM_HOST));
				if (err == null) err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_COUNTERS.invokeExact(handle, pinCnt));
				if (err == null) err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_FEC_COUNTERS.invokeExact(
					handle, pinCnt.asSlice(32, 8), pinCnt.asSlice(40, 8)));
				// frames: int32 n | int32 chan[4] | int64 bit_index[4] | int32 errors[4] | uint8 data[4][256]
				MemorySegment fN = pinFrames.asSlice(0, 4), fChan = pinFrames.asSlice(8, 16), fAt = pinFrames.asSlice(24, 32),
					fErr = pinFrames.asSlice(56, 16), fData = pinFrames.asSlice(72, 1024);
				if (err == null) err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_FRAMES.invokeExact(
					handle, fN, fChan, fAt, fErr, fData, MAX_FRAMES));
				if (err == null && doFFT) err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_CENTRE.invokeExact(handle, pinMisc.asSlice(8, 4)));
				if (err != null) { fail(err); return false; }
				MemorySegment.copy(pinDs, ValueLayout.JAVA_DOUBLE, 0, ds, 0, 2 * nDs);
				MemorySegment.copy(pinDm, ValueLayout.JAVA_DOUBLE, 0, dm, 0, 2 * nDs);
				nBits = Math.min(pinNbits.get(ValueLayout.JAVA_INT, 0), maxBits);
				MemorySegment.copy(pinBits, ValueLayout.JAVA_BYTE, 0, bits, 0, nBits);
				MemorySegment.copy(pinBitAt, ValueLayout.JAVA_LONG, 0, bitAt, 0, nBits);
				cntRaw = pinCnt.getAtIndex(ValueLayout.JAVA_LONG, 0);
				cntDS = pinCnt.getAtIndex(ValueLayout.JAVA_LONG, 1);
				cntBit = pinCnt.getAtIndex(ValueLayout.JAVA_LONG, 2);
				cntFEC = pinCnt.getAtIndex(ValueLayout.JAVA_LONG, 4);
				cntDec = pinCnt.getAtIndex(ValueLayout.JAVA_LONG, 5);
				firstDs = cntDS - nDs;
				int nf = Math.min(fN.get(ValueLayout.JAVA_INT, 0), MAX_FRAMES);
				frameErrors = Integer.MIN_VALUE;
				if (nf > 0) {                                                        // the last frame of the block (:565-569)
					frameErrors = fErr.getAtIndex(ValueLayout.JAVA_INT, nf - 1);
					MemorySegment.copy(fData, ValueLayout.JAVA_BYTE, 256L * (nf - 1), frame, 0, 256);
				}
				if (doFFT) centreBin = pinMisc.get(ValueLayout.JAVA_INT, 8);
				return true;
			} catch (Throwable t) {
				fail("CudaFUNcubeBPSKDemod: " + t);
				return false;
			}
		}
	}

	private void fail(String msg) {
		failed = true;                                                               // stay on the Java path from now on
		logger.statusMsg(msg);
	}

	/** Caller holds ctx.lock. */
	private void close() {
		try {
			if (handle != null) { int rc = (int) JsdrCuda.BPSK_DESTROY.invokeExact(handle); }
		} catch (Throwable t) { }
		handle = null;
		for (MemorySegment p : new MemorySegment[] {pinIn, pinDs, pinDm, pinBits, pinBitAt, pinNbits, pinCnt, pinFrames, pinMisc})
			ctx.free(p);
		pinIn = pinDs = pinDm = pinBits = pinBitAt = pinNbits = pinCnt = pinFrames = pinMisc = null;
	}
}
