// CudaFUNcubeBPSKDemod.java — drop-in for the DSP of N FUNcubeBPSKDemod instances
// (jsdr.java:479-483 builds "jsdr-funcube-demods" of them): one bank of tuners on
// the GPU fed by the same receive(buf).  Same config keys and published keys as
// FUNcubeBPSKDemod.java:97-99,195-200,377-378.  The bit stream, which the reference
// never publishes (SURVEY Q11), goes to the reference's own FECDecoder on the Java
// side exactly as FUNcubeBPSKDemod.java:553-574 does.
//
// NOT COMPILED HERE (no JDK in the build image); see INTEGRATION.md.
package com.ashbysoft.java_sdr;

import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

public class CudaFUNcubeBPSKDemod implements IAudioHandler, IPublishListener {
	private static final int FEC_BITS_SIZE = 5200, SYNC_VECTOR_SIZE = 65;
	private final IConfig config;
	private final IPublish publish;
	private final ILogger logger;
	private final JsdrCuda.Context ctx;
	private final int nchan;
	private AudioDescriptor adsc;
	private MemorySegment handle, pinIn, pinBits, pinNbits, pinTuning;
	private int samples, maxBits;
	private final byte[][] fecCorr;          // dmFECCorr per tuner (:503)
	private final FECDecoder[] decoders;
	private final byte[][] decoded;
	private final byte[] syncVector;

	public CudaFUNcubeBPSKDemod(int count, IConfig cfg, IPublish pub, ILogger log, IUIHost hst, IAudio aud,
		JsdrCuda.Context ctx, byte[] syncVector) {
		this.nchan = count;
		this.config = cfg;
		this.publish = pub;
		this.logger = log;
		this.ctx = ctx;
		this.syncVector = syncVector;         // FUNcubeBPSKDemod.SYNC_VECTOR (:79-81)
		fecCorr = new byte[count][FEC_BITS_SIZE];
		decoded = new byte[count][256];
		decoders = new FECDecoder[count];
		for (int i = 0; i < count; i++) decoders[i] = new FECDecoder();
		setup(aud);
		pub.listen(this);
	}

	public void notify(String key, Object val) {
		if ("audio-change".equals(key) && val instanceof IAudio) setup((IAudio) val);   // :167-171
	}

	private synchronized void setup(IAudio audio) {
		try {
			if (handle != null) { int rc = (int) JsdrCuda.BPSK_DESTROY.invokeExact(handle); }
			adsc = audio.getAudioDescriptor();
			samples = adsc.blen / adsc.size;                                         // :194
			pinTuning = ctx.pinned(8L * nchan);
			for (int i = 0; i < nchan; i++)
				pinTuning.setAtIndex(ValueLayout.JAVA_DOUBLE, i,
					(double) config.getIntConfig("FUNcube" + i + "-bpsk-tuning", 12000));  // :195
			MemorySegment out = ctx.arena.allocate(ValueLayout.ADDRESS);
			String err = JsdrCuda.check((int) JsdrCuda.BPSK_CREATE.invokeExact(
				ctx.handle, adsc.rate, nchan, pinTuning, samples, out));
			if (err != null) { logger.statusMsg(err); handle = null; return; }
			handle = out.get(ValueLayout.ADDRESS, 0);
			maxBits = samples * 9600 / adsc.rate + 2;
			pinIn = ctx.pinned(8L * samples);
			pinBits = ctx.pinned((long) nchan * maxBits);
			pinNbits = ctx.pinned(4L * nchan);
			audio.remHandler(this);                                                  // :207-208
			audio.addHandler(this);
		} catch (Throwable t) {
			logger.statusMsg("CudaFUNcubeBPSKDemod setup: " + t);
			handle = null;
		}
	}

	public void setTuning(int chan, double hz) {                                     // actionPerformed :174-189
		try {
			String err = JsdrCuda.check((int) JsdrCuda.BPSK_SET_TUNING.invokeExact(handle, chan, hz));
			if (err != null) logger.statusMsg(err);
			config.setIntConfig("FUNcube" + chan + "-bpsk-tuning", (int) hz);
		} catch (Throwable t) {
			logger.statusMsg("CudaFUNcubeBPSKDemod: " + t);
		}
	}

	public synchronized void receive(float[] buf) {                                  // :358-379
		if (handle == null) return;
		try {
			MemorySegment.copy(buf, 0, pinIn, ValueLayout.JAVA_FLOAT, 0, 2 * samples);
			// chan_stride 0: one stream fans out to every tuner, as in jsdr.java:479-483
			String err = JsdrCuda.check((int) JsdrCuda.BPSK_RECEIVE_F32.invokeExact(
				handle, pinIn, samples, 0L, JsdrCuda.MEM_HOST));
			if (err == null)
				err = JsdrCuda.check((int) JsdrCuda.BPSK_READ_BITS.invokeExact(
					handle, pinBits, MemorySegment.NULL, pinNbits, maxBits, JsdrCuda.MEM_HOST));
			if (err != null) { logger.statusMsg(err); return; }
			for (int c = 0; c < nchan; c++) {
				int nb = pinNbits.getAtIndex(ValueLayout.JAVA_INT, c);
				for (int k = 0; k < nb && k < maxBits; k++)
					pushBit(c, pinBits.get(ValueLayout.JAVA_BYTE, (long) c * maxBits + k));
				publish.setPublish("FUNcube" + c + "-bpsk-centre", -1);              // :377
				publish.setPublish("FUNcube" + c + "-bpsk-tune",
					config.getIntConfig("FUNcube" + c + "-bpsk-tuning", 12000));     // :378
			}
		} catch (Throwable t) {
			logger.statusMsg("CudaFUNcubeBPSKDemod: " + t);
		}
	}

	// FUNcubeBPSKDemod.java:553-574 unchanged: rolling symbol buffer, sync correlation, FEC
	private void pushBit(int c, byte bit) {
		byte[] corr = fecCorr[c];
		System.arraycopy(corr, 1, corr, 0, corr.length - 1);
		corr[corr.length - 1] = bit;
		int dmCorr = 0;
		for (int n = 0; n < SYNC_VECTOR_SIZE; n++) dmCorr += corr[n * 80] * syncVector[n];
		if (dmCorr >= 45) {
			byte[] fecBits = new byte[FEC_BITS_SIZE];
			for (int n = 0; n < FEC_BITS_SIZE; n++) fecBits[n] = (byte) (corr[n] == 1 ? 0xc0 : 0x40);
			int errs = decoders[c].FECDecode(fecBits, decoded[c]);
			if (errs >= 0) publish.setPublish("FUNcube" + c + "-bpsk-frame", decoded[c].clone());
		}
	}
}
