"""TEST INFRASTRUCTURE ONLY (never imported by the product): the reference's sensor synchronisation restated with
the same pandas calls, line for line (nexar_videos.py:318-341), minus the file / cv2 access.  Parity pinning: the product
path it checks is ALSO held to the sensor rows the unmodified ``_load_and_sync_sensor_data`` returned for 756 cases
(tests/golden/dataset_sensor.npz, tests/test_dataset_golden.py)."""
import numpy as np
import pandas as pd


def sync_sensor_pandas(time_sec, accel, frame_count, fps):
    df_sensor = pd.DataFrame({"time_sec": np.asarray(time_sec, dtype=np.float64),
                              "accel_x_G": accel[:, 0], "accel_y_G": accel[:, 1],
                              "accel_z_G": accel[:, 2], "accel_total_G": accel[:, 3]})
    sensor_start_time = df_sensor["time_sec"].iloc[0]                                   # :324
    df_sensor["relative_time_sec"] = df_sensor["time_sec"] - sensor_start_time          # :325
    df_sensor = df_sensor.set_index("relative_time_sec")                                # :326
    video_times = pd.Series([i / fps for i in range(frame_count)], name="video_time_sec")   # :329
    aligned = df_sensor.reindex(df_sensor.index.union(video_times)).interpolate("index").loc[video_times]  # :332
    aligned.reset_index(drop=True, inplace=True)
    return aligned[["accel_x_G", "accel_y_G", "accel_z_G", "accel_total_G"]].values    # :340-341
