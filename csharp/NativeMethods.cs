// NativeMethods.cs -- P/Invoke declarations for libmmlb200.so, 1:1 with include/mmlb200.h.
// Add to src/MyMediaLite/ (MyMediaLite.csproj) next to the three Cuda* classes in this directory.
// Not compiled in this repository's CI: the image has no .NET/Mono toolchain (see INTEGRATION.md).
using System;
using System.Runtime.InteropServices;

namespace MyMediaLite.Native
{
	/// <summary>mml_mf_params of include/mmlb200.h (sequential layout, 4-byte fields)</summary>
	[StructLayout(LayoutKind.Sequential)]
	public struct MmlMfParams
	{
		public int biased, num_factors;
		public float learn_rate, decay, regularization, bias_learn_rate, bias_reg, reg_u, reg_i;
		public int frequency_regularization, loss, bold_driver, max_threads;
		public int schedule, num_groups, num_subgroups, group_rule, persistent;
		public float hot_item_factor;
		public int hot_copies, intra_block, hot_merge_average, async_workers, ctas_per_group;
	}

	/// <summary>mml_wrmf_params of include/mmlb200.h</summary>
	[StructLayout(LayoutKind.Sequential)]
	public struct MmlWrmfParams
	{
		public int num_factors;
		public double alpha, regularization;
	}

	/// <summary>Ref-counted owner of a native handle: Recommender.Clone() is a MemberwiseClone (Recommender.cs:37-40),
	/// so clones alias the handle; Train() always allocates a fresh one instead of mutating a shared one.</summary>
	public sealed class MmlHandle : SafeHandle
	{
		readonly Func<IntPtr, int> destroy;
		public MmlHandle(IntPtr h, Func<IntPtr, int> destroy) : base(IntPtr.Zero, true) { SetHandle(h); this.destroy = destroy; }
		public override bool IsInvalid { get { return handle == IntPtr.Zero; } }
		protected override bool ReleaseHandle() { return destroy(handle) == 0; }
	}

	public static class Mml
	{
		const string LIB = "mmlb200";   // libmmlb200.so on the library path

		public const int LOSS_RMSE = 0, LOSS_MAE = 1, LOSS_LOGISTIC = 2;
		public const int SCHEDULE_SERIAL = 0, SCHEDULE_DSGD = 1, SCHEDULE_NAIVE = 2;
		public const int TOPN_AUTO = 0, TOPN_EXACT = 1, TOPN_TENSOR = 2;

		[DllImport(LIB)] static extern IntPtr mml_last_error();
		[DllImport(LIB)] static extern IntPtr mml_version();
		public static string LastError() { return Marshal.PtrToStringAnsi(mml_last_error()); }
		public static string Version() { return Marshal.PtrToStringAnsi(mml_version()); }
		/// <summary>non-zero status -> managed exception carrying mml_last_error() (SURVEY.md section 8b, "Errors")</summary>
		public static void Check(int status) { if (status != 0) throw new InvalidOperationException(LastError()); }

		// context
		[DllImport(LIB)] public static extern int mml_ctx_create(int n_gpus, int[] device_ids, out IntPtr ctx);
		[DllImport(LIB)] public static extern int mml_dist_unique_id(byte[] out128);
		[DllImport(LIB)] public static extern int mml_ctx_create_dist(int rank, int world, int device, byte[] unique_id128, out IntPtr ctx);
		[DllImport(LIB)] public static extern int mml_ctx_destroy(IntPtr ctx);
		[DllImport(LIB)] public static extern int mml_ctx_synchronize(IntPtr ctx);
		[DllImport(LIB)] public static extern int mml_ctx_flush_l2(IntPtr ctx);
		[DllImport(LIB)] public static extern int mml_ctx_probe_l2(IntPtr ctx, int mode, int n_rows, int row_floats, int reps, out double rows_per_s);
		[DllImport(LIB)] public static extern int mml_ctx_sm_count(IntPtr ctx, out int sm_count);

		// ingest (text file -> id mapping -> COO in pinned host memory)
		public const int FILE_RATINGS = 0, FILE_RATINGS_NO_VALUE = 1, FILE_FEEDBACK = 2;
		public const int MAP_IDENTITY = 0, MAP_FIRST_SEEN = 1;
		public const int ERR_FORMAT = 6, ERR_IO = 7;
		[DllImport(LIB)] public static extern int mml_ingest_file(string path, int kind, int user_mapping, int item_mapping, int ignore_first_line, int n_threads, IntPtr prior, out IntPtr ingest);
		[DllImport(LIB)] public static extern int mml_ingest_text(byte[] text, long len, int kind, int user_mapping, int item_mapping, int ignore_first_line, int n_threads, IntPtr prior, out IntPtr ingest);
		[DllImport(LIB)] public static extern int mml_ingest_destroy(IntPtr ingest);
		[DllImport(LIB)] public static extern int mml_ingest_info(IntPtr ingest, out long n, out int max_user, out int max_item, out int n_user_ids, out int n_item_ids, out int pinned);
		[DllImport(LIB)] public static extern int mml_ingest_copy(IntPtr ingest, [Out] int[] users, [Out] int[] items, [Out] float[] values);
		[DllImport(LIB)] public static extern int mml_ingest_original_ids(IntPtr ingest, int which, int first, int count, [Out] byte[] buf, long buf_len, [Out] long[] offsets, out long needed);
		[DllImport(LIB)] public static extern int mml_ingest_to_ratings(IntPtr ctx, IntPtr ingest, out IntPtr ratings);
		[DllImport(LIB)] public static extern int mml_ingest_to_feedback(IntPtr ctx, IntPtr ingest, out IntPtr feedback);
		/// <summary>MML_ERR_FORMAT -> FormatException with the reference's message, MML_ERR_IO -> IOException</summary>
		public static void CheckIngest(int status)
		{
			if (status == ERR_FORMAT) throw new FormatException(LastError());
			if (status == ERR_IO) throw new System.IO.IOException(LastError());
			Check(status);
		}

		// rating matrix build
		[DllImport(LIB)] public static extern int mml_ratings_create(IntPtr ctx, int[] users, int[] items, float[] values, long n, int max_user, int max_item, out IntPtr ratings);
		[DllImport(LIB)] public static extern int mml_ratings_destroy(IntPtr ratings);
		[DllImport(LIB)] public static extern int mml_ratings_counts(IntPtr ratings, int by_item, [Out] int[] counts);
		[DllImport(LIB)] public static extern int mml_ratings_csr(IntPtr ratings, int by_item, [Out] long[] row_ptr, [Out] int[] idx);
		[DllImport(LIB)] public static extern int mml_ratings_stats(IntPtr ratings, out float average, out float min_rating, out float max_rating);
		[DllImport(LIB)] public static extern int mml_shuffle_apply(IntPtr ctx, [In, Out] int[] perm, int[] H, long n);
		[DllImport(LIB)] public static extern int mml_partition_blocks(IntPtr ratings, int[] user_perm, int[] item_perm, int g, [Out] long[] block_ptr, [Out] int[] idx);
		[DllImport(LIB)] public static extern int mml_partition_indices(IntPtr ctx, int[] random_index, long n, int num_groups, [Out] long[] list_ptr, [Out] int[] idx);

		// MatrixFactorization / BiasedMatrixFactorization
		[DllImport(LIB)] public static extern void mml_mf_params_default(out MmlMfParams p);
		[DllImport(LIB)] public static extern int mml_sgd_create(IntPtr ctx, IntPtr ratings, ref MmlMfParams p, int[] user_perm, int[] item_perm, out IntPtr model);
		[DllImport(LIB)] public static extern int mml_sgd_destroy(IntPtr model);
		[DllImport(LIB)] public static extern int mml_sgd_set_model(IntPtr model, float[] user_factors, float[] item_factors, float[] user_bias, float[] item_bias);
		[DllImport(LIB)] public static extern int mml_sgd_init_model(IntPtr model, ulong seed, double init_mean, double init_stddev);
		[DllImport(LIB)] public static extern int mml_sgd_get_model(IntPtr model, [Out] float[] user_factors, [Out] float[] item_factors, [Out] float[] user_bias, [Out] float[] item_bias, out float global_bias, out float current_learnrate);
		[DllImport(LIB)] public static extern int mml_sgd_set_learnrate(IntPtr model, float current_learnrate);
		[DllImport(LIB)] public static extern int mml_sgd_set_scale(IntPtr model, float min_rating, float max_rating, float global_bias);
		[DllImport(LIB)] public static extern int mml_sgd_iterate(IntPtr model, int[] subepoch_sequence, int[] random_index, long n_index);
		[DllImport(LIB)] public static extern int mml_sgd_invalidate_index(IntPtr model);
		[DllImport(LIB)] public static extern int mml_sgd_iterate_indices(IntPtr model, int[] indices, long n, int update_user, int update_item);
		[DllImport(LIB)] public static extern int mml_sgd_learn_factors(IntPtr model, int[] indices, long n, int update_user, int update_item, int num_iter);
		[DllImport(LIB)] public static extern int mml_sgd_predict(IntPtr model, int[] users, int[] items, long n, [Out] float[] result);
		[DllImport(LIB)] public static extern int mml_sgd_fold_in(IntPtr model, long[] rated_ptr, int[] rated_items, float[] rated_values, long n_users, float[] init_factors, int num_iter, [Out] float[] out_vectors);
		[DllImport(LIB)] public static extern int mml_sgd_score_items(IntPtr model, float[] user_vectors, long n_users, int[] candidates, long n_cand, [Out] float[] out_scores);
		[DllImport(LIB)] public static extern int mml_sgd_set_rows(IntPtr model, int by_item, int[] ids, long n, float[] factors, float[] biases);
		[DllImport(LIB)] public static extern int mml_sgd_evaluate(IntPtr model, int[] users, int[] items, float[] values, long n, [Out] float[] out4);
		[DllImport(LIB)] public static extern int mml_sgd_evaluate_train(IntPtr model, [Out] float[] out4);
		[DllImport(LIB)] public static extern int mml_sgd_objective(IntPtr model, out double objective);
		[DllImport(LIB)] public static extern int mml_sgd_stats(IntPtr model, out long kernel_launches, out float last_iterate_ms);
		[DllImport(LIB)] public static extern int mml_sgd_strata_info(IntPtr model, out int G, out int W, out long n_rounds, out long staged_bytes);
		[DllImport(LIB)] public static extern int mml_sgd_hot_items(IntPtr model, out long n_hot);
		[DllImport(LIB)] public static extern int mml_sgd_grid(IntPtr model, out int G, out int ctas_per_group);
		[DllImport(LIB)] public static extern int mml_sgd_schedule_dump(IntPtr model, int[] subepoch_sequence, [Out] int[] order, [Out] int[] block, [Out] int[] copy, [Out] int[] round);

		// top-N
		[DllImport(LIB)] public static extern int mml_topn_mf(IntPtr ctx, float[] user_factors, int n_model_users, float[] item_factors, int n_model_items, int k,
			int[] users, long n_users, int n, int[] candidates, long n_cand, long[] ignore_ptr, int[] ignore_idx,
			[Out] int[] out_items, [Out] float[] out_scores, [Out] int[] out_counts);
		[DllImport(LIB)] public static extern int mml_items_evaluate_mf(IntPtr ctx, float[] user_factors, int n_model_users, float[] item_factors, int n_model_items, int k,
			int[] test_users, long n_test_users, int[] candidates, long n_cand, long[] test_ptr, int[] test_idx, long[] ignore_ptr, int[] ignore_idx, int n,
			[Out] float[] out_measures, [Out] int[] out_used);
		[DllImport(LIB)] public static extern int mml_topn_set_mode(int mode);
		[DllImport(LIB)] public static extern int mml_topn_set_filter(int kind);
		[DllImport(LIB)] public static extern int mml_topn_last_stats(out long users_tensor_path, out long users_exact_path, out float tensor_path_ms);

		// WRMF
		[DllImport(LIB)] public static extern int mml_feedback_create(IntPtr ctx, int[] users, int[] items, long n, int max_user, int max_item, out IntPtr feedback);
		[DllImport(LIB)] public static extern int mml_feedback_destroy(IntPtr feedback);
		[DllImport(LIB)] public static extern int mml_feedback_nnz(IntPtr feedback, out long nnz);
		[DllImport(LIB)] public static extern int mml_feedback_csr(IntPtr feedback, int by_item, [Out] long[] row_ptr, [Out] int[] cols);
		[DllImport(LIB)] public static extern int mml_wrmf_create(IntPtr ctx, IntPtr feedback, ref MmlWrmfParams p, out IntPtr model);
		[DllImport(LIB)] public static extern int mml_wrmf_destroy(IntPtr model);
		[DllImport(LIB)] public static extern int mml_wrmf_set_model(IntPtr model, float[] user_factors, float[] item_factors);
		[DllImport(LIB)] public static extern int mml_wrmf_init_model(IntPtr model, ulong seed, double init_mean, double init_stddev);
		[DllImport(LIB)] public static extern int mml_wrmf_get_model(IntPtr model, [Out] float[] user_factors, [Out] float[] item_factors);
		[DllImport(LIB)] public static extern int mml_wrmf_iterate(IntPtr model);
		[DllImport(LIB)] public static extern int mml_wrmf_retrain(IntPtr model, int by_item, int[] ids, long n);
		[DllImport(LIB)] public static extern int mml_wrmf_set_mode(int mode);
		[DllImport(LIB)] public static extern int mml_wrmf_debug_gram(IntPtr model, [Out] float[] out_gram, out int out_user);
		[DllImport(LIB)] public static extern int mml_wrmf_evaluate(IntPtr model, int[] test_users, long n_test_users, int[] candidates, long n_cand,
			long[] test_ptr, int[] test_idx, long[] ignore_ptr, int[] ignore_idx, int n, [Out] float[] out_measures, [Out] int[] out_used);
		[DllImport(LIB)] public static extern int mml_wrmf_stats(IntPtr model, out long kernel_launches, out float last_iterate_ms);
		[DllImport(LIB)] public static extern int mml_wrmf_shard(IntPtr model, int by_item, [Out] int[] ranges);
		[DllImport(LIB)] public static extern int mml_wrmf_recommend(IntPtr model, int[] users, long n_users, int n, int[] candidates, long n_cand, long[] ignore_ptr, int[] ignore_idx,
			[Out] int[] out_items, [Out] float[] out_scores, [Out] int[] out_counts);

		static readonly object ctx_lock = new object();
		static readonly Dictionary<uint, IntPtr> shared_ctx = new Dictionary<uint, IntPtr>();
		/// <summary>One library context per process and GPU count: GPUs 0 .. n_gpus - 1, all driven from this process
		/// (mml_ctx_create with n_gpus > 1 = the NumGpus property). Throws without a CUDA device: there is no CPU path.</summary>
		public static IntPtr Context(uint n_gpus = 1)
		{
			lock (ctx_lock)
			{
				if (n_gpus < 1) n_gpus = 1;
				IntPtr ctx;
				if (!shared_ctx.TryGetValue(n_gpus, out ctx))
				{
					Check(mml_ctx_create((int) n_gpus, null, out ctx));
					shared_ctx[n_gpus] = ctx;
				}
				return ctx;
			}
		}

		/// <summary>The one engine knob (process-wide; the option set of the classes stays the reference's + NumGpus).
		/// Auto (default): Iterate() runs the parallel epoch kernel whatever MaxThreads says (MaxThreads keeps its other meaning,
		/// UpdateLearnRate twice per epoch when > 1), except on data sets below SerialBelow ratings, where the exact
		/// single-threaded order costs nothing. Reference: MaxThreads = 1 walks RandomIndex in the reference's order on one warp
		/// (parity runs). Parallel: always the parallel kernel. Environment: MMLB200_ORDER = auto | reference | parallel.</summary>
		public enum EngineOrder { Auto, Reference, Parallel }
		public const int SerialBelow = 20000;
		public static EngineOrder Order = ParseOrder(Environment.GetEnvironmentVariable("MMLB200_ORDER"));
		static EngineOrder ParseOrder(string v)
		{
			return v == "reference" ? EngineOrder.Reference : (v == "parallel" ? EngineOrder.Parallel : EngineOrder.Auto);
		}
	}
}
